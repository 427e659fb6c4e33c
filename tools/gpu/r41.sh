#!/bin/bash
# round 2, call 2: training check with TF32 off on the reference side, precision probe of the post-VQ decoders, GPU suite after the backend refactor
mkdir -p gpurun_out
timeout 600 python tests/flow_train_gpu_check.py > gpurun_out/r41_train_check.log 2>&1; echo "train check rc=$?"; grep -c "^ok" gpurun_out/r41_train_check.log; grep FAIL gpurun_out/r41_train_check.log
timeout 900 python tools/gpu/precision_probe.py > gpurun_out/r41_precision_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r41_precision_probe.txt | tail -30
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r41_pytest_gpu.log 2>&1; tail -5 gpurun_out/r41_pytest_gpu.log
