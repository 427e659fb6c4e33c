#!/bin/bash
# mode 4 (bf16x3) bring-up: tests, full-size parity of every mode against the cuDNN fp32 path, benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_dcn_gpu.py -m gpu -q --tb=line 2>&1 | tail -30 > gpurun_out/r17_conv_tests.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r17_pipe_tests.log
timeout 400 python - > gpurun_out/r17_fullsize_parity.log 2>&1 <<'PY'
import torch, sys, math
sys.path.insert(0, ".")
from glare_b200 import synth
from glare_b200.dense import make_dense
from glare_b200.engine import GlareEngine
sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
lq, gt = synth.synth_images(2, 400, 600, seed=0)
lr = synth.preprocess(synth.pad_lol(lq))
res = {}
for name in ("torch-fp32", "tc-3xtf32", "tc-tf32bf16x2", "tc-bf16x3", "tc-tf32", "tc-bf16"):
    eng = GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense(name))
    st = {}
    out = eng.infer(lr, stages=st)
    res[name] = (out.float().cpu(), st["idx"].cpu(), st["z_flow"].float().cpu(), st["cond_feat"].float().cpu())
    del eng
    torch.cuda.empty_cache()
ref = res["torch-fp32"]
mse = lambda a: float(((a[:, :, :400, 20:].clamp(0, 1) - gt) ** 2).mean())
for name, (o, idx, z, cf) in res.items():
    print("%-16s idx agree %.5f  cond_feat maxdiff %.3g  z maxdiff %.3g  pixel maxdiff %.3g  mean abs %.3g  dPSNR %.5f dB" % (
        name, float((idx == ref[1]).float().mean()), float((cf - ref[3]).abs().max()), float((z - ref[2]).abs().max()),
        float((o - ref[0]).abs().max()), float((o - ref[0]).abs().mean()), 10 * math.log10(mse(ref[0]) / mse(o))))
PY
for d in tc-bf16x3 tc-tf32bf16x2; do
  timeout 600 python bench.py --steps 4 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r17_bench_$d.json 2> gpurun_out/r17_bench_$d.err
done
tail -4 gpurun_out/r17_conv_tests.log | cut -c1-200; tail -4 gpurun_out/r17_pipe_tests.log | cut -c1-200; tail -7 gpurun_out/r17_fullsize_parity.log; cut -c1-220 gpurun_out/r17_bench_tc-bf16x3.json; tail -n 2 gpurun_out/r17_bench_tc-bf16x3.err
