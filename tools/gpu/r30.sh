#!/bin/bash
# dcn_tc: ring depth vs L1 for the corner gathers (A/B 6 / 4 / 3 / 2 stages)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dcn_gpu.py tests/test_conv_tc_gpu.py -m gpu -q --tb=short -k "dcn or attention" 2>&1 | tail -15 > gpurun_out/r30_tests.log
grep -E "passed|failed|error" gpurun_out/r30_tests.log | tail -3
if grep -q "failed\|error" gpurun_out/r30_tests.log; then cat gpurun_out/r30_tests.log; fi
for n in 6 4 3 2; do
GLARE_DCN_STAGES=$n timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r30_bench_s$n.json 2> gpurun_out/r30_bench_s$n.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r30_bench_s$n.json").read().strip().splitlines()[-1])
    print("dcn stages $n", d["value"], d["ms_per_step"], "dcn_tc", d["breakdown_ms_per_step"]["dcn_tc"], "attn_softmax", d["breakdown_ms_per_step"]["attn_softmax"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r30_bench_s$n.err").read()[-2000:])
PY
done
