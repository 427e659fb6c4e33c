#!/bin/bash
# fused-softmax attention, branch-free exp epilogue: parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=short -k "attention" 2>&1 | tail -30 > gpurun_out/r23_tests_attn.log
grep -E "passed|failed|error" gpurun_out/r23_tests_attn.log | tail -3
if grep -q "failed\|error" gpurun_out/r23_tests_attn.log; then cat gpurun_out/r23_tests_attn.log; fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r23_bench_default.json 2> gpurun_out/r23_bench_default.err
for f in default; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r23_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["breakdown_ms_per_step"], d["config"].get("library_fallbacks_per_run"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r23_bench_$f.err").read()[-2000:])
PY
done
