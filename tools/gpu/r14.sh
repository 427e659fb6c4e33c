#!/bin/bash
# per-warp independent TMA-store epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r14_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r14_bench_default.json 2> gpurun_out/r14_bench_default.err
timeout 600 python bench.py --steps 4 --warmup 3 --dense tc-bf16 --no-cpu-baseline > gpurun_out/r14_bench_tc-bf16.json 2> gpurun_out/r14_bench_tc-bf16.err
grep -E "passed|failed" gpurun_out/r14_tests.log | tail -3; cut -c1-250 gpurun_out/r14_bench_default.json; tail -n 3 gpurun_out/r14_bench_default.err
