#!/bin/bash
# ninth GPU call: ncu --set full on the default (mode 3) conv_tc launches at batch 15 to find the binding resource per layer type
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -s 2 -c 60 -o gpurun_out/r9_prof_mode3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu.log 2>&1
tail -3 gpurun_out/r9_ncu.log | cut -c1-300
