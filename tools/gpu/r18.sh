#!/bin/bash
# full validation of the default (bf16x3, CTA pairs): all GPU tests, smoke, bench (+cpu baseline), reference arm, traffic, launch list, ncu full (CSV)
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r18_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r18_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r18_bench_default.json 2> gpurun_out/r18_bench_default.err
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' -s 1600 -c 500 --csv --log-file gpurun_out/r18_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r18_ncu_traffic.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r18_ncu_launches.log 2>&1
python tools/ncu_summary.py launches /tmp/ncu/launches.csv > gpurun_out/r18_launches_summary.txt 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:'conv_tc_kernel' -s 2 -c 30 -o /tmp/ncu/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r18_ncu_full.log 2>&1
ncu -i /tmp/ncu/prof.ncu-rep --page raw --csv > gpurun_out/r18_conv_tc_full_raw.csv 2>/dev/null
grep -E "passed|failed" gpurun_out/r18_tests.log | tail -2; tail -1 gpurun_out/r18_smoke.log; cut -c1-300 gpurun_out/r18_bench_default.json; head -12 gpurun_out/r18_launches_summary.txt
