#!/bin/bash
# round 2, call 1: stage-2 training kernels' first hardware run (log kept), RING2 A/B, train probe
mkdir -p gpurun_out
timeout 600 python tests/flow_train_gpu_check.py > gpurun_out/r40_train_check.log 2>&1; echo "train check rc=$?"; tail -40 gpurun_out/r40_train_check.log
timeout 300 python tests/conv_ring2_gpu_check.py > gpurun_out/r40_ring2_default.log 2>&1; echo "ring2 default rc=$?"; tail -5 gpurun_out/r40_ring2_default.log
GLARE_CONV_RING2=1 timeout 300 python tests/conv_ring2_gpu_check.py > gpurun_out/r40_ring2_on.log 2>&1; echo "ring2 on rc=$?"; tail -5 gpurun_out/r40_ring2_on.log
timeout 900 python tools/gpu/train_probe.py 3 > gpurun_out/r40_train_probe.txt 2>&1; echo "probe rc=$?"; tail -8 gpurun_out/r40_train_probe.txt
GLARE_WGRAD_TC=1 timeout 900 python tools/gpu/train_probe.py 3 > gpurun_out/r40_train_probe_wgrad_tc.txt 2>&1; echo "probe tc rc=$?"; tail -8 gpurun_out/r40_train_probe_wgrad_tc.txt
