#!/bin/bash
# eighth GPU call: 2-CTA clusters with multicast weight tiles, sub-pixel upsample conv, weight-cache fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_dcn_gpu.py -m gpu -q --tb=line 2>&1 | tail -40 > gpurun_out/r8_conv_tests.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_api_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r8_pipe_tests.log
for d in tc-tf32bf16x2 tc-3xtf32 tc-bf16; do
  timeout 600 python bench.py --steps 3 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r8_bench_$d.json 2> gpurun_out/r8_bench_$d.err
done
GLARE_CONV_NO_CLUSTER=1 timeout 600 python bench.py --steps 3 --warmup 3 --dense tc-tf32bf16x2 --no-cpu-baseline > gpurun_out/r8_bench_tc-tf32bf16x2_nocluster.json 2> gpurun_out/r8_bench_tc-tf32bf16x2_nocluster.err
GLARE_CONV_NO_CLUSTER=1 timeout 600 python bench.py --steps 3 --warmup 3 --dense tc-bf16 --no-cpu-baseline > gpurun_out/r8_bench_tc-bf16_nocluster.json 2> gpurun_out/r8_bench_tc-bf16_nocluster.err
tail -5 gpurun_out/r8_conv_tests.log; tail -5 gpurun_out/r8_pipe_tests.log; cat gpurun_out/r8_bench_*.json | cut -c1-200; tail -n 2 gpurun_out/r8_bench_*.err
