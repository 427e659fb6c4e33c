#!/bin/bash
# tenth GPU call: CTA-pair (cta_group::2) variant of conv_tc, A/B against the multicast-cluster variant
mkdir -p gpurun_out
GLARE_CONV_PAIR=1 timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=line -x 2>&1 | tail -30 > gpurun_out/r10_conv_tests_pair.log
GLARE_CONV_PAIR=1 timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r10_pipe_tests_pair.log
for d in tc-tf32bf16x2 tc-bf16; do
  GLARE_CONV_PAIR=1 timeout 600 python bench.py --steps 3 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r10_bench_${d}_pair.json 2> gpurun_out/r10_bench_${d}_pair.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench_default_mcast.json 2> gpurun_out/r10_bench_default_mcast.err
tail -4 gpurun_out/r10_conv_tests_pair.log; tail -4 gpurun_out/r10_pipe_tests_pair.log; cat gpurun_out/r10_bench_*.json | cut -c1-200; tail -n 2 gpurun_out/r10_bench_*.err
