"""Golden vector at the 1920x1080 shape (BASELINE.json configs[4]; infer_unpaired.py auto_padding -> 1088x1936, 131 648 latent tokens).
TEST INFRASTRUCTURE; authoring container, ~40 min of CPU:   python -m oracle.gen_golden_1080p   -> tests/golden/pipe_1080p.npz

The REFERENCE cannot run at this size on this host (its AttnBlock materialises a 69 GB score matrix per sample, 11 times), so the
vector comes from the ORACLE -- pinned to the reference at every smaller size, 420x620 included (PIN_REPORT.txt) -- with the attention
evaluated in blocks of query rows (same arithmetic per row; 5e-7 from the one-shot form where both fit).  Stored: z (flow output), the VQ
indices, the RGB output at stride 2 in fp16, and the oracle's PSNR against the synthetic ground truth."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glare_b200 import synth  # noqa: E402
from oracle import glare_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
    lq, gt = synth.synth_images(1, 1080, 1920, seed=200)
    xp, (h1, h2, w1, w2) = synth.auto_padding(lq)
    lr = synth.preprocess(xp)
    st = {}
    t0 = time.perf_counter()
    out = O.glare_infer(sd_g, sd_v, lr, stages=st)
    dt = time.perf_counter() - t0
    crop = out[:, :, h1:out.shape[2] - h2, w1:out.shape[3] - w2].clamp(0, 1)
    psnr = O.psnr(crop, gt)
    msg = ("pipe_1080p (1920x1080 -> %dx%d, %d tokens): ORACLE only (reference needs 69 GB per attention matrix), %.0f s on %d threads; "
           "|z| max %.3g, |out| max %.3g, PSNR vs synthetic gt %.4f dB" %
           (lr.shape[2], lr.shape[3], st["idx"].numel(), dt, torch.get_num_threads(), float(st["z_flow"].abs().max()), float(out.abs().max()), psnr))
    print(msg)
    np.savez_compressed(os.path.join(GOLD, "pipe_1080p.npz"), z_flow=st["z_flow"].numpy(), idx=st["idx"].reshape(-1).numpy().astype(np.int16),
                        out_s2=out[:, :, ::2, ::2].numpy().astype(np.float16), psnr=np.float64(psnr), pad=np.array([h1, h2, w1, w2]),
                        lr_checksum=np.float64(lr.double().sum().item()), seconds=np.float64(dt))
    rep = os.path.join(GOLD, "PIN_REPORT.txt")
    lines = [l for l in open(rep).read().splitlines() if not l.startswith("pipe_1080p")]
    with open(rep, "w") as f:
        f.write("\n".join(lines + [msg]) + "\n")


if __name__ == "__main__":
    main()
