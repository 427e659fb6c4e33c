#!/bin/bash
# band test, 1080p single image (fused vs exact softmax), launch list + conv_tc DRAM traffic of the default configuration
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=short -k "attention" 2>&1 | tail -15 > gpurun_out/r28_tests_attn.log
grep -E "passed|failed|error" gpurun_out/r28_tests_attn.log | tail -3
if grep -q "failed\|error" gpurun_out/r28_tests_attn.log; then cat gpurun_out/r28_tests_attn.log; fi
timeout 900 python - > gpurun_out/r28_1080p.log 2>&1 <<'PY'
import sys, time, math, torch
sys.path.insert(0, ".")
from glare_b200 import synth
from glare_b200.api import GlareEnhancer
from glare_b200.dense import make_dense
sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
lq, gt = synth.synth_images(1, 1080, 1920, seed=1)
u8 = (lq.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
outs = {}
for name in ("fused", "exact"):
    dense = make_dense("auto")
    dense.attn_fused = name == "fused"
    enh = GlareEnhancer(sd_g, sd_v, device="cuda:0", pad="auto", dense=dense)
    enh.enhance(u8)                                   # warm-up (weight packing, allocator)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = enh.enhance(u8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    outs[name] = out.float() / 255.0
    print("%-6s softmax, 1080x1920 (padded 1088x1936, 131648 latent tokens), bf16x3: %.3f s/image end to end, peak memory %.1f GB, finite %s, mean %.4f, fallbacks %s"
          % (name, dt, torch.cuda.max_memory_allocated() / 2**30, bool(torch.isfinite(outs[name]).all()), float(outs[name].mean()), dense.fallbacks))
    del enh
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
mse = float(((outs["fused"] - outs["exact"]) ** 2).mean())
print("fused vs exact (uint8 outputs): PSNR %.2f dB, max abs diff %.4f, differing pixels %.5f %%" % (10 * math.log10(1.0 / max(mse, 1e-12)), float((outs["fused"] - outs["exact"]).abs().max()), 100 * float((outs["fused"] != outs["exact"]).float().mean())))
PY
cat gpurun_out/r28_1080p.log | tail -4
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r28_ncu_launches.log 2>&1
python tools/ncu_summary.py launches /tmp/ncu/launches.csv > gpurun_out/r28_launches_summary.txt 2>&1
python tools/traffic_summary.py /tmp/ncu/launches.csv > gpurun_out/r28_traffic.json 2>&1
head -40 gpurun_out/r28_launches_summary.txt
