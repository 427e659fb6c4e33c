#!/bin/bash
# DCN-TC sampler: offsets/mask tile staged in shared memory (A/B against the global-load path)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dcn_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -20 > gpurun_out/r21_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_default.json 2> gpurun_out/r21_bench_default.err
GLARE_DCN_NO_OM_SMEM=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_noomsmem.json 2> gpurun_out/r21_bench_noomsmem.err
grep -E "passed|failed" gpurun_out/r21_tests.log | tail -2
for f in default noomsmem; do python - <<PY
import json
d=json.loads(open("gpurun_out/r21_bench_$f.json").read().strip().splitlines()[-1])
print("$f", d["value"], d["ms_per_step"], d["roofline"].get("breakdown_ms_per_step"))
PY
done
