"""Golden vector for the OTHER branch of the stage-2 objective: `mean = gt` (LLFlowVQGAN2_arch.py:109, taken with probability
train_gt_ratio = 0.2 under train_stage2_LOL.yml), from the unmodified reference's autograd on the inputs of tests/golden/stage2.npz.
TEST INFRASTRUCTURE; authoring container only.   python -m oracle.gen_golden_stage2_gtmean   -> tests/golden/stage2_gtmean.npz"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glare_b200 import synth  # noqa: E402
from oracle import glare_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(8)
    with np.load(os.path.join(GOLD, "stage2.npz")) as f:
        gt_lat, lr_img = torch.from_numpy(f["gt_latent"]), torch.from_numpy(f["lr"])
    n2, _, opt = ref_shims.build_reference_stage2(seed=0)
    sd2 = synth.synth_state_dict("netG_stage2", 0)
    n2.load_state_dict(sd2, strict=True)
    n2.train()
    for m in n2.modules():
        if hasattr(m, "inited"):
            m.inited = True
    ratio = opt["train_gt_ratio"]
    random.seed(1)
    while True:                       # next draw that takes the gt branch: not (random.random() > train_gt_ratio)
        st = random.getstate()
        if not (random.random() > ratio):
            random.setstate(st)
            break
    z2, nll_ref, _ = n2(gt=gt_lat, lr=lr_img, reverse=False)
    nll_ref.mean().backward()
    sd2g = {k: v.clone().requires_grad_(True) for k, v in sd2.items()}
    z2o, nll_o = O.stage2_nll(sd2g, gt_lat, lr_img, use_gt_mean=True)
    nll_o.mean().backward()
    named = dict(n2.named_parameters())
    msg = ["stage2 gt-mean branch (train_gt_ratio %.2f): nll ref %s oracle %s ; z maxdiff %.3g" %
           (ratio, nll_ref.tolist(), nll_o.tolist(), float((z2o - z2).abs().max()))]
    gsel = {}
    for k in ("flowUpsamplerNet.layers.0.actnorm.bias", "flowUpsamplerNet.layers.2.affine.fAffine.0.weight",
              "flowUpsamplerNet.layers.27.affine.fAffine.4.weight", "RRDB.cond_conv.0.weight", "RRDB.encoder.conv_in.weight"):
        gr, go = named[k].grad, sd2g[k].grad
        msg.append("  grad %-55s |g| %.3g  oracle-vs-ref maxdiff %.3g" % (k, float(gr.abs().max()), float((gr - go).abs().max())))
        gsel["grad." + k] = gr.numpy()
    no_grad = [k for k in ("RRDB.color_conv.weight", "RRDB.color_conv.bias") if named[k].grad is None or float(named[k].grad.abs().max()) == 0.0]
    msg.append("  parameters without gradient on this branch: %s" % no_grad)
    print("\n".join(msg))
    np.savez_compressed(os.path.join(GOLD, "stage2_gtmean.npz"), nll=nll_ref.detach().numpy(), z=z2.detach().numpy(),
                        no_grad=np.array(no_grad), **gsel)
    rep = os.path.join(GOLD, "PIN_REPORT.txt")
    lines = [l for l in open(rep).read().splitlines() if "gt-mean" not in l and not l.startswith("  grad* ")]
    with open(rep, "w") as f:
        f.write("\n".join(lines + [m.replace("  grad ", "  grad* ") for m in msg]) + "\n")


if __name__ == "__main__":
    main()
