#!/bin/bash
# fused-softmax attention (exp in the scores-GEMM epilogue, 1/rowsum in the P V epilogue): parity, then A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=short -k "attention" 2>&1 | tail -30 > gpurun_out/r22_tests_attn.log
grep -E "passed|failed|error" gpurun_out/r22_tests_attn.log | tail -3
if grep -q "failed\|error" gpurun_out/r22_tests_attn.log; then cat gpurun_out/r22_tests_attn.log; fi
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_api_gpu.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r22_tests_pipe.log
grep -E "passed|failed|error" gpurun_out/r22_tests_pipe.log | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r22_bench_default.json 2> gpurun_out/r22_bench_default.err
GLARE_ATTN_UNFUSED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r22_bench_unfused.json 2> gpurun_out/r22_bench_unfused.err
for f in default unfused; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r22_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["breakdown_ms_per_step"], d["config"].get("library_fallbacks_per_run"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r22_bench_$f.err").read()[-2000:])
PY
done
