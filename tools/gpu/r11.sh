#!/bin/bash
# eleventh GPU call: full validation of the default configuration (mode tf32+2xbf16, CTA pairs): all GPU tests, smoke, bench (+cpu baseline),
# reference arm, DRAM traffic of conv_tc, launch list of one step
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r11_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r11_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r11_bench_default.json 2> gpurun_out/r11_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r11_bench_reference.json 2> gpurun_out/r11_bench_reference.err
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' -s 1600 -c 500 --csv --log-file gpurun_out/r11_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r11_ncu_traffic.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r11_ncu_launches.log 2>&1
python tools/ncu_summary.py launches /tmp/ncu/launches.csv > gpurun_out/r11_launches_summary.txt 2>&1
grep -E "passed|failed" gpurun_out/r11_tests.log | tail -2; tail -1 gpurun_out/r11_smoke.log; cut -c1-300 gpurun_out/r11_bench_default.json; head -12 gpurun_out/r11_launches_summary.txt
