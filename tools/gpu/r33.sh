#!/bin/bash
# operand-emitting conv epilogues (q / k / offset-conv / attention output): parity + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_pipeline_gpu.py tests/test_api_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/r33_tests.log
grep -E "passed|failed|error" gpurun_out/r33_tests.log | tail -3
if grep -q "failed\|error" gpurun_out/r33_tests.log; then cat gpurun_out/r33_tests.log; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r33_bench_pack.json 2> gpurun_out/r33_bench_pack.err
GLARE_NO_PACK_EPILOGUE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r33_bench_nopack.json 2> gpurun_out/r33_bench_nopack.err
for f in pack nopack; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r33_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["breakdown_ms_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r33_bench_$f.err").read()[-2000:])
PY
done
