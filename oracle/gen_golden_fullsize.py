"""Golden vector of the whole inference path at the BENCHMARK shape (BASELINE.json configs[1]: 400x600 reflect-padded to 420x620), produced
by the UNMODIFIED reference modules.  TEST INFRASTRUCTURE; run in the authoring container only (needs /root/reference, ~2 min of CPU):

    python -m oracle.gen_golden_fullsize      # writes tests/golden/pipe_420x620.npz, appends to tests/golden/PIN_REPORT.txt

Stored: the reference's z (flow output), VQ indices and RGB output for image 0 of synth.synth_images(1, 400, 600, seed=0) -- the first image
of the bench batch -- plus the oracle-vs-reference differences at this size.  The input is regenerated from the seed by the tests
(`lr_checksum` guards against a synth change); encoder features are too large to commit (133 MB), the teacher-forced test feeds the golden z.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glare_b200 import synth  # noqa: E402
from oracle import glare_oracle as O  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    netG, net_hq, _ = ref_shims.build_reference("LOL.yml", seed=0)
    sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
    netG.load_state_dict(sd_g, strict=True)
    net_hq.load_state_dict(sd_v, strict=True)
    netG.eval(), net_hq.eval()
    lq, gt = synth.synth_images(1, 400, 600, seed=0)
    lr = synth.preprocess(synth.pad_lol(lq))
    cap = {}
    hooks = [netG.RRDB.register_forward_hook(lambda m, i, o: cap.__setitem__("enc", o)),
             net_hq.quantize.register_forward_hook(lambda m, i, o: cap.__setitem__("vq", (i[0], o)))]
    t0 = time.perf_counter()
    with torch.no_grad():
        out_ref, z_ref = netG(net_vq=net_hq, lr=lr, reverse=True)
    t_ref = time.perf_counter() - t0
    for hk in hooks:
        hk.remove()
    idx_ref = cap["vq"][1][2][2].reshape(-1)
    st = {}
    t0 = time.perf_counter()
    out_o = O.glare_infer(sd_g, sd_v, lr, stages=st)
    t_o = time.perf_counter() - t0
    md = lambda a, b: float((a.double() - b.double()).abs().max())          # noqa: E731
    agree = float((st["idx"].reshape(-1) == idx_ref).float().mean())
    msg = ("pipe_420x620 (bench shape, image 0 of the batch): cond_feat %.3g color_map %.3g z_flow %.3g (|z| %.3g) idx agree %.5f out %.3g "
           "(|out| %.3g); reference %.1f s, oracle %.1f s on %d threads" %
           (md(st["cond_feat"], cap["enc"]["cond_feat"]), md(st["color_map"], cap["enc"]["color_map"]), md(st["z_flow"], z_ref),
            float(z_ref.abs().max()), agree, md(out_o, out_ref), float(out_ref.abs().max()), t_ref, t_o, torch.get_num_threads()))
    print(msg)
    np.savez_compressed(os.path.join(GOLD, "pipe_420x620.npz"), z_flow=z_ref.numpy(), idx=idx_ref.numpy().astype(np.int16),
                        color_map=cap["enc"]["color_map"].numpy(), out=out_ref.numpy(),
                        lr_checksum=np.float64(lr.double().sum().item()),
                        fingerprint_netG=np.float64(synth.state_fingerprint(sd_g)), fingerprint_vqgan=np.float64(synth.state_fingerprint(sd_v)))
    rep = os.path.join(GOLD, "PIN_REPORT.txt")
    lines = [l for l in open(rep).read().splitlines() if not l.startswith("pipe_420x620")]
    with open(rep, "w") as f:
        f.write("\n".join(lines + [msg]) + "\n")


if __name__ == "__main__":
    main()
