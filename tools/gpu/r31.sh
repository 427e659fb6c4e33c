#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 300 python tools/gpu/dcn_probe.py 2 > gpurun_out/r31_dcn_probe.txt 2>&1
cat gpurun_out/r31_dcn_probe.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'dcn_tc_kernel' -s 2 -c 1 -o /tmp/ncu/dcn python tools/gpu/dcn_probe.py 2 > gpurun_out/r31_ncu.log 2>&1
ncu -i /tmp/ncu/dcn.ncu-rep --page raw --csv > gpurun_out/r31_dcn_raw.csv 2>/dev/null
ncu -i /tmp/ncu/dcn.ncu-rep --page details > gpurun_out/r31_dcn_details.txt 2>/dev/null
ls -la gpurun_out/r31_dcn_raw.csv gpurun_out/r31_dcn_details.txt
