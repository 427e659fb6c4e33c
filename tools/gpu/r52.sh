#!/bin/bash
# round 2, call 13: skinny weight-gradient kernel: training tests + kernel-time breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_flow_train_gpu.py tests/test_modules_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/gpu/train_probe.py 3 --profile > gpurun_out/r52_train_probe_profile.txt 2>&1; grep -A14 "stage-2 step" gpurun_out/r52_train_probe_profile.txt | head -18
