#!/bin/bash
# round 2, call 10: ncu evidence for the current code: conv_tc DRAM traffic of one step (feeds roofline.traffic, tagged with the source hash),
# --set full on the first conv_tc launches of a step, launch list of the whole bench command
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-alt --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' -s 1410 -c 470 --csv --log-file /tmp/ncu/traffic.csv $CMD > gpurun_out/r49_ncu_traffic.log 2>&1; echo "traffic rc=$?"
python tools/traffic_summary.py /tmp/ncu/traffic.csv > gpurun_out/r49_traffic_conv_tc.json 2>&1
python tools/traffic_summary.py /tmp/ncu/traffic.csv --bench-json tcgen05-bf16x3 15 "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 1410 -c 470 $CMD (tools/gpu/r49.sh)" > gpurun_out/conv_tc_traffic.json 2>&1; cat gpurun_out/conv_tc_traffic.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -s 1410 -c 90 -o /tmp/ncu/prof $CMD > gpurun_out/r49_ncu_full.log 2>&1; echo "full rc=$?"
ncu -i /tmp/ncu/prof.ncu-rep --page raw --csv > gpurun_out/r49_conv_tc_full_raw.csv 2>/dev/null; ls -la /tmp/ncu/
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv $CMD > gpurun_out/r49_ncu_launches.log 2>&1; echo "launches rc=$?"
python tools/ncu_summary.py launches /tmp/ncu/launches.csv > gpurun_out/r49_launches_summary.txt 2>&1; head -30 gpurun_out/r49_launches_summary.txt
# one more: dcn_tc + gn kernels full set (second and third kernels by time)
timeout 600 ncu --set full --clock-control none -k regex:'dcn_tc_kernel|gn_apply_kernel|flow_tail_kernel|vq_argmin' -s 20 -c 12 -o /tmp/ncu/prof2 $CMD > gpurun_out/r49_ncu_full2.log 2>&1
ncu -i /tmp/ncu/prof2.ncu-rep --page raw --csv > gpurun_out/r49_other_full_raw.csv 2>/dev/null
ls -la gpurun_out | grep r49
