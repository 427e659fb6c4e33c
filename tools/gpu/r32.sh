#!/bin/bash
# dcn_tc: 16 sampler warps (two groups of eight, 2 pixel rows x 4 corners in flight per lane) vs 8 (A/B), new row mapping, F2FP split
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dcn_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r32_tests.log
grep -E "passed|failed|error" gpurun_out/r32_tests.log | tail -3
if grep -q "failed\|error" gpurun_out/r32_tests.log; then cat gpurun_out/r32_tests.log; fi
timeout 300 python tools/gpu/dcn_probe.py 2
GLARE_DCN_SW=4 timeout 300 python tools/gpu/dcn_probe.py 2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r32_bench_sw8.json 2> gpurun_out/r32_bench_sw8.err
GLARE_DCN_SW=4 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r32_bench_sw4.json 2> gpurun_out/r32_bench_sw4.err
for f in sw8 sw4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r32_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], "dcn_tc", d["breakdown_ms_per_step"]["dcn_tc"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r32_bench_$f.err").read()[-2000:])
PY
done
