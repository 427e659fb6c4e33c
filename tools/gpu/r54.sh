#!/bin/bash
# round 2, call 15: conv_tc DRAM traffic of one step re-captured for the final kernel sources (hash-tagged), full GPU suite, smoke
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-alt --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' -s 1410 -c 470 --csv --log-file /tmp/ncu/traffic.csv $CMD > gpurun_out/r54_ncu_traffic.log 2>&1; echo "traffic rc=$?"
python tools/traffic_summary.py /tmp/ncu/traffic.csv > gpurun_out/r54_traffic_conv_tc.json 2>&1
python tools/traffic_summary.py /tmp/ncu/traffic.csv --bench-json tcgen05-bf16x3 15 "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 1410 -c 470 $CMD (tools/gpu/r54.sh)" > gpurun_out/conv_tc_traffic.json 2>&1; cat gpurun_out/conv_tc_traffic.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r54_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r54_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r54_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r54_smoke.log
