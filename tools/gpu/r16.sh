#!/bin/bash
# API tests with the pre/post kernels, batch-1 latency (CPU launch bound?), bf16x... sanity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_api_gpu.py -m gpu -q --tb=short 2>&1 | tail -20 > gpurun_out/r16_api_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/r16_bench_b1.json 2> gpurun_out/r16_bench_b1.err
timeout 600 python bench.py --steps 5 --warmup 3 --batch 4 --no-cpu-baseline > gpurun_out/r16_bench_b4.json 2> gpurun_out/r16_bench_b4.err
tail -3 gpurun_out/r16_api_tests.log; cut -c1-260 gpurun_out/r16_bench_b1.json; cut -c1-260 gpurun_out/r16_bench_b4.json
