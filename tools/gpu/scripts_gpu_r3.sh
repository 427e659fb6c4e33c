#!/bin/bash
# third GPU call: re-verify conv after tile-order / 3-D weight-map change, attention GEMM path, breakdown benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=line 2>&1 | tail -40 > gpurun_out/r3_conv_tests.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r3_pipe_tests.log
for d in tc-3xtf32 tc-3xtf32+libattn tc-bf16 tc-bf16+libattn; do
  timeout 600 python bench.py --steps 2 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r3_bench_$d.json 2> gpurun_out/r3_bench_$d.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r3_launches_bf16.csv python bench.py --steps 1 --warmup 3 --batch 1 --dense tc-bf16 --no-cpu-baseline > gpurun_out/r3_ncu_bf16.log 2>&1
tail -5 gpurun_out/r3_conv_tests.log; tail -5 gpurun_out/r3_pipe_tests.log; cat gpurun_out/r3_bench_*.json | cut -c1-400; tail -n 3 gpurun_out/r3_bench_*.err
