"""The dense operators through the vendor libraries (cuDNN convolutions, cuBLAS attention matmuls, ATen GroupNorm / softmax; tests/libdense.py,
test infrastructure) inside the same engine -- VQ, flow and DCN stay glare kernels -- next to the tcgen05 path, batch 15 x 420x620:
what the reference's nn.Conv2d / AttnBlock / Normalize calls cost on a B200 when recompiled, per precision.  python tools/gpu/library_baseline.py"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from glare_b200 import synth  # noqa: E402
from glare_b200.dense import make_dense  # noqa: E402
from glare_b200.engine import GlareEngine  # noqa: E402
from libdense import TorchDense  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 15
sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
lq, gt = synth.synth_images(B, 400, 600, seed=0)
lr = synth.preprocess(synth.pad_lol(lq)).cuda()


def timed(eng, n=3):
    with torch.no_grad():
        eng.infer(lr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = eng.infer(lr)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ref_ms, ref = timed(GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense("auto")), 5)
print("batch %d x 420x620" % B)
print("  glare tcgen05 bf16x3 (fp32-grade)            %8.1f ms/step  %6.2f images/s" % (ref_ms, B / ref_ms * 1e3))
for name, dense in (("cuDNN / cuBLAS fp32 (TF32 off)", TorchDense(torch.float32, allow_tf32=False)),
                    ("cuDNN / cuBLAS fp32 with TF32", TorchDense(torch.float32, allow_tf32=True)),
                    ("cuDNN / cuBLAS bf16", TorchDense(torch.bfloat16))):
    ms, out = timed(GlareEngine(sd_g, sd_v, device="cuda:0", dense=dense))
    d = (out.float() - ref.float()).abs()
    print("  %-44s %8.1f ms/step  %6.2f images/s   (x%.1f slower; pixel mean abs diff vs glare %.2e)" % (name, ms, B / ms * 1e3, ms / ref_ms, float(d.mean())))
ms16, _ = timed(GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense("tc-bf16")), 5)
print("  glare tcgen05 bf16 operands                  %8.1f ms/step  %6.2f images/s" % (ms16, B / ms16 * 1e3))
