#!/bin/bash
# 8-GPU weak-scaling check of the bench exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r19_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r19_bench_8gpu.json 2> gpurun_out/r19_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > gpurun_out/r19_bench_8gpu_reference.json 2> gpurun_out/r19_bench_8gpu_reference.err
cut -c1-300 gpurun_out/r19_bench_8gpu.json; tail -n 3 gpurun_out/r19_bench_8gpu.err; cut -c1-200 gpurun_out/r19_bench_8gpu_reference.json
