#!/bin/bash
# seventh GPU call: mode 3 (tf32 + 2x bf16 cross terms) bring-up and comparison with 3xTF32
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_dcn_gpu.py -m gpu -q --tb=line 2>&1 | tail -40 > gpurun_out/r7_conv_tests.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short -rP 2>&1 | tail -60 > gpurun_out/r7_pipe_tests.log
for d in tc-tf32bf16x2 tc-3xtf32; do
  timeout 600 python bench.py --steps 3 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r7_bench_$d.json 2> gpurun_out/r7_bench_$d.err
done
timeout 300 python - > gpurun_out/r7_fullsize_parity.log 2>&1 <<'PY'
# full-size (420x620) agreement between the fp32-grade tensor-core modes and the cuDNN fp32 library path (no oracle at this size in seconds)
import torch, sys
sys.path.insert(0, ".")
from glare_b200 import synth
from glare_b200.dense import make_dense
from glare_b200.engine import GlareEngine
sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
lq, gt = synth.synth_images(2, 400, 600, seed=0)
lr = synth.preprocess(synth.pad_lol(lq))
res = {}
for name in ("torch-fp32", "tc-3xtf32", "tc-tf32bf16x2", "tc-tf32", "tc-bf16"):
    eng = GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense(name))
    st = {}
    out = eng.infer(lr, stages=st)
    res[name] = (out.float().cpu(), st["idx"].cpu(), st["z_flow"].float().cpu())
    del eng
    torch.cuda.empty_cache()
ref = res["torch-fp32"]
for name, (o, idx, z) in res.items():
    mse = lambda a: float(((a[:, :, :400, 20:].clamp(0, 1) - gt) ** 2).mean())
    import math
    print("%-16s idx agree %.5f  z maxdiff %.3g  pixel maxdiff %.3g  mean abs %.3g  dPSNR %.5f dB" % (
        name, float((idx == ref[1]).float().mean()), float((z - ref[2]).abs().max()), float((o - ref[0]).abs().max()),
        float((o - ref[0]).abs().mean()), 10 * math.log10(mse(ref[0]) / mse(o))))
PY
tail -4 gpurun_out/r7_conv_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r7_pipe_tests.log | tail -5; cat gpurun_out/r7_fullsize_parity.log | tail -8; cat gpurun_out/r7_bench_*.json | cut -c1-200
