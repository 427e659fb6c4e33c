#!/bin/bash
# fifth GPU call (2 GPUs): all-tcgen05 dense path (no cuDNN), reworked DCN sampler, 1- and 2-GPU benches
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES=0,1
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=line 2>&1 | tail -30 > gpurun_out/r5_conv_tests.log
timeout 900 python -m pytest tests/test_dcn_gpu.py -m gpu -q --tb=line 2>&1 | tail -30 > gpurun_out/r5_dcn_tests.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r5_pipe_tests.log
for d in tc-3xtf32 tc-bf16; do
  timeout 600 python bench.py --steps 2 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r5_bench_$d.json 2> gpurun_out/r5_bench_$d.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench_2gpu.json 2> gpurun_out/r5_bench_2gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r5_bench_reference.json 2> gpurun_out/r5_bench_reference.err
tail -4 gpurun_out/r5_conv_tests.log; tail -6 gpurun_out/r5_dcn_tests.log; tail -6 gpurun_out/r5_pipe_tests.log; cat gpurun_out/r5_bench_*.json | cut -c1-250; tail -n 3 gpurun_out/r5_bench_2gpu.err
