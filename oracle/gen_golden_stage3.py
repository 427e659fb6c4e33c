"""Golden vector for one stage-3 training evaluation (train_stage3.py / VQLLFLOWDModel.optimize_parameters, VQLLFLOWD_model.py:187-232) from
the UNMODIFIED reference on CPU: the reference's generator in train mode called with reverse=True, reverse_with_grad=True, its own MS-SSIM
function, its PerceptualNetwork forward code over torchvision's vgg16.features[:16] structure with seeded weights (the pretrained download is
not available), the three-term objective, and torch autograd for the gradients of every `deformable_decoder.*` parameter.
TEST INFRASTRUCTURE; authoring container only.   python -m oracle.gen_golden_stage3   -> tests/golden/stage3.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glare_b200 import losses as GL  # noqa: E402  (layer table only)
from glare_b200 import synth  # noqa: E402
from oracle import glare_oracle as O  # noqa: E402
from oracle import losses as OL  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SELECT = ("deformable_decoder.conv_in.weight", "deformable_decoder.mid.attn_1.q.bias", "deformable_decoder.mid.attn_1.norm.weight", "deformable_decoder.up.0.block.2.conv2.weight",
          "deformable_decoder.up.1.block.0.nin_shortcut.weight", "deformable_decoder.up.1.upsample.conv.bias", "deformable_decoder.mix.0.w",
          "deformable_decoder.mix.1.w", "deformable_decoder.warp.0.offset.bias", "deformable_decoder.warp.0.dcn.conv_offset.bias",
          "deformable_decoder.warp.1.dcn.bias", "deformable_decoder.warp.1.dcn.conv_offset.weight", "deformable_decoder.norm_out.weight",
          "deformable_decoder.residual_conv.weight", "deformable_decoder.residual_conv.bias")


def vgg_state(seed=0):
    g = torch.Generator().manual_seed(4000 + seed)
    sd = {}
    for idx, ci, co in GL.VGG_CONVS:
        sd["%d.weight" % idx] = torch.randn((co, ci, 3, 3), generator=g) * (2.0 / (9 * ci)) ** 0.5
        sd["%d.bias" % idx] = 0.05 * torch.randn((co,), generator=g)
    return sd


def main():
    torch.set_num_threads(8)
    netG, net_hq, _ = ref_shims.build_reference("LOL.yml", seed=0)
    sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
    netG.load_state_dict(sd_g, strict=True)
    net_hq.load_state_dict(sd_v, strict=True)
    netG.train()
    lq, gt = synth.synth_images(2, 32, 48, seed=5)                       # batch 2 like train_stage3_LOL.yml:39 (the :567 ratio spans the batch)
    lr = synth.preprocess(lq)
    from models.modules.pytorch_msssim import msssim as ref_msssim
    import models.modules.losses as ref_losses
    from torchvision.models import vgg16
    vsd = vgg_state(0)
    percep = ref_losses.RefPerceptualNetwork.__new__(ref_losses.RefPerceptualNetwork)
    torch.nn.Module.__init__(percep)
    percep.vgg_model = vgg16(weights=None).features[:16]
    percep.vgg_model.load_state_dict(vsd, strict=True)
    percep.layer_name_mapping = {'3': "relu1_2", '8': "relu2_2", '15': "relu3_3"}

    rec, enc_feat = netG(net_vq=net_hq, lr=lr, reverse=True, reverse_with_grad=True, epses=None, lr_enc=None)
    rec = rec.to(torch.float32)
    sr = rec.clamp(0, 1)                                                  # VQLLFLOWD_model.py:212-223, verbatim semantics
    not_nan = ~torch.isnan(sr)
    sr[torch.isnan(sr)] = 0
    terms = {"l1_loss": ((sr - gt) * not_nan).abs().mean(), "percep_loss": percep(sr, gt) * 0.01,
             "ssim_loss": (1 - ref_msssim(sr, gt, normalize=True)) * 0.2}
    total = sum(terms.values())
    total.backward()
    named = dict(netG.named_parameters())
    with_grad = sorted(k for k, p in named.items() if p.grad is not None)
    assert all(k.startswith("deformable_decoder.") for k in with_grad), [k for k in with_grad if not k.startswith("deformable_decoder.")][:3]

    # the oracle on the same inputs (frozen stages + decoder + losses), autograd through a differentiable DCN
    from torchvision.ops import deform_conv2d
    dcn = lambda x, off, m, w, b, **kw: deform_conv2d(x, off, w, b, padding=1, mask=m)        # noqa: E731
    with torch.no_grad():
        st = {}
        O.glare_infer(sd_g, sd_v, lr, per_sample_ratio=False, stages=st)
    sda = {k: v.clone().requires_grad_(True) for k, v in sd_g.items() if k.startswith("deformable_decoder.")}
    rec_o = O.aft_decoder(sda, st["z_flow"], [st["vq_feat1"], st["vq_feat0"]], {1: st["mid1"], 0: st["mid0"]}, per_sample_ratio=False, dcn=dcn)
    total_o, terms_o = OL.stage3_loss(rec_o, gt, vsd)
    total_o.backward()
    msg = ["stage3 (batch 2 x 32x48): total ref %.8f oracle %.8f ; rec maxdiff %.3g ; terms ref %s" %
           (float(total), float(total_o), float((rec_o - rec).abs().max()), {k: round(float(v), 8) for k, v in terms.items()})]
    gmax = max(float(named[k].grad.abs().max()) for k in with_grad)
    worst, wk = 0.0, None
    for k in with_grad:
        gr, go = named[k].grad, sda[k].grad
        rel = float((gr - go).abs().max()) / max(float(gr.abs().max()), 1e-6 * gmax)       # k.bias gradients are 0 in exact arithmetic
        if rel > worst:
            worst, wk = rel, k
    msg.append("  stage3 gradients of %d deformable_decoder parameters: worst oracle-vs-ref relative maxdiff %.3g (%s)" % (len(with_grad), worst, wk))
    print("\n".join(msg))
    out = {"lq": lq.numpy(), "gt": gt.numpy(), "rec": rec.detach().numpy(), "z_flow": enc_feat.detach().numpy(),
           "total": np.float64(float(total)), "with_grad": np.array(with_grad),
           "abs_sum": np.array([float(named[k].grad.double().abs().sum()) for k in with_grad])}
    out.update({"term." + k: np.float64(float(v)) for k, v in terms.items()})
    out.update({"grad." + k: named[k].grad.numpy() for k in SELECT})
    np.savez_compressed(os.path.join(GOLD, "stage3.npz"), **out)
    rep = os.path.join(GOLD, "PIN_REPORT.txt")
    lines = [l for l in open(rep).read().splitlines() if "stage3" not in l]
    with open(rep, "w") as f:
        f.write("\n".join(lines + msg) + "\n")


if __name__ == "__main__":
    main()
