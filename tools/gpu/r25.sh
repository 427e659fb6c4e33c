#!/bin/bash
# epilogue: double-buffered TMEM loads, full-chunk fast path, ex2-based fused softmax numerator: full GPU suite + probe + bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 > gpurun_out/r25_tests.log
grep -E "passed|failed|error" gpurun_out/r25_tests.log | tail -3
if grep -q "failed\|error" gpurun_out/r25_tests.log; then cat gpurun_out/r25_tests.log; fi
timeout 300 python tools/gpu/attn_probe.py 10 > gpurun_out/r25_attn_probe.txt 2>&1
cat gpurun_out/r25_attn_probe.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r25_bench_default.json 2> gpurun_out/r25_bench_default.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r25_bench_default.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["breakdown_ms_per_step"], d["config"].get("library_fallbacks_per_run"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r25_bench_default.err").read()[-2000:])
PY
