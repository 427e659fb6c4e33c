#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 300 python tools/gpu/attn_probe.py 10 > gpurun_out/r24_attn_probe.txt 2>&1
cat gpurun_out/r24_attn_probe.txt
timeout 600 ncu --set full --clock-control none -k regex:'conv_tc_kernel' -c 24 -o /tmp/ncu/attn python tools/gpu/attn_probe.py 1 > gpurun_out/r24_ncu.log 2>&1
ncu -i /tmp/ncu/attn.ncu-rep --page raw --csv > gpurun_out/r24_attn_raw.csv 2>/dev/null
ls -la gpurun_out/r24_attn_raw.csv
