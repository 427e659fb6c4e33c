#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/flow_train_gpu_check.py > gpurun_out/r64_train_check.log 2>&1; echo "check rc=$?"; grep -E "FAIL|wgrad_conv|stage-2 step" gpurun_out/r64_train_check.log | head -16
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_zz_flow_train_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/gpu/train_probe.py 3 --profile > gpurun_out/r64_train_probe_profile.txt 2>&1; grep -A12 "stage-2 step" gpurun_out/r64_train_probe_profile.txt | head -16
