#!/bin/bash
# round 2, call 5 (2 GPUs): bench under torchrun at N=2 (graph + async gather path, alt configs incl. the stage-2 step with its all-reduce), N=1 DCN compare
mkdir -p gpurun_out
timeout 600 python tools/gpu/dcn_ref_compare.py 15 > gpurun_out/r44_dcn_ref_compare.txt 2>&1; echo "dcn ref rc=$?"; tail -14 gpurun_out/r44_dcn_ref_compare.txt
timeout 600 python -m pytest tests/test_dcn_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r44_bench_2gpu.json 2> gpurun_out/r44_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 5000 gpurun_out/r44_bench_2gpu.json; tail -5 gpurun_out/r44_bench_2gpu.err
