#!/bin/bash
# GroupNorm kernels with more loads in flight; input validation; re-run of the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_pipeline_gpu.py tests/test_api_gpu.py -m gpu -q --tb=short 2>&1 | tail -20 > gpurun_out/r20_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r20_bench_default.json 2> gpurun_out/r20_bench_default.err
grep -E "passed|failed" gpurun_out/r20_tests.log | tail -2; cut -c1-250 gpurun_out/r20_bench_default.json; tail -n 2 gpurun_out/r20_bench_default.err
