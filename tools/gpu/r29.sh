#!/bin/bash
# filter-row halo staging for BN <= 128 convs, faster GroupNorm kernels, fused AFT elementwise kernel: parity + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/r29_tests_conv.log
grep -E "passed|failed|error" gpurun_out/r29_tests_conv.log | tail -3
if grep -q "failed\|error" gpurun_out/r29_tests_conv.log; then cat gpurun_out/r29_tests_conv.log; fi
timeout 900 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_conv_tc_gpu.py 2>&1 | tail -25 > gpurun_out/r29_tests_rest.log
grep -E "passed|failed|error" gpurun_out/r29_tests_rest.log | tail -3
if grep -q "failed\|error" gpurun_out/r29_tests_rest.log; then cat gpurun_out/r29_tests_rest.log; fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_default.json 2> gpurun_out/r29_bench_default.err
GLARE_CONV_NO_HALO=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r29_bench_nohalo.json 2> gpurun_out/r29_bench_nohalo.err
for f in default nohalo; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r29_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["breakdown_ms_per_step"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r29_bench_$f.err").read()[-2000:])
PY
done
