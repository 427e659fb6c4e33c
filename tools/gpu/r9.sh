#!/bin/bash
# ncu --set full on the default (mode 3) conv_tc launches at batch 15; only the CSV pages travel back (the .ncu-rep is too big)
mkdir -p gpurun_out /tmp/ncu
timeout 1500 ncu --set full --clock-control none -k regex:'conv_tc_kernel' -s 2 -c 40 -o /tmp/ncu/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r9_ncu.log 2>&1
ncu -i /tmp/ncu/prof.ncu-rep --page raw --csv > gpurun_out/r9_mode3_raw.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out/r9_mode3_raw.csv
