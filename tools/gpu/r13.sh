#!/bin/bash
# thirteenth GPU call: BASELINE config 5 shape (1920x1080 unpaired, auto padding to 1088x1936) through the public API, default and bf16
mkdir -p gpurun_out
timeout 900 python - > gpurun_out/r13_1080p.log 2>&1 <<'PY'
import sys, time, math, torch
sys.path.insert(0, ".")
from glare_b200 import synth
from glare_b200.api import GlareEnhancer
from glare_b200.dense import make_dense
sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
lq, gt = synth.synth_images(1, 1080, 1920, seed=1)
u8 = (lq.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
outs = {}
for name in ("auto", "tc-bf16"):
    enh = GlareEnhancer(sd_g, sd_v, device="cuda:0", pad="auto", dense=make_dense(name))
    enh.enhance(u8)                                   # warm-up (weight packing, allocator)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = enh.enhance(u8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    outs[name] = out.float() / 255.0
    print("%-8s 1080x1920 (padded 1088x1936, 131648 latent tokens): %.3f s/image end to end, peak memory %.1f GB, output finite %s, mean %.4f"
          % (name, dt, torch.cuda.max_memory_allocated() / 2**30, bool(torch.isfinite(outs[name]).all()), float(outs[name].mean())))
    del enh
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
mse = float(((outs["auto"] - outs["tc-bf16"]) ** 2).mean())
print("bf16 vs fp32-grade: PSNR between outputs %.2f dB, max abs diff %.4f" % (10 * math.log10(1.0 / max(mse, 1e-12)), float((outs["auto"] - outs["tc-bf16"]).abs().max())))
PY
cat gpurun_out/r13_1080p.log | tail -5
