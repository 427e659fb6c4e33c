#!/bin/bash
# final validation of the round-1 default (bf16x3, CTA pairs, fused softmax, halo staging, operand epilogues): all GPU tests, smoke, bench (+cpu baseline),
# reference arm, conv_tc DRAM traffic, ncu --set full on the first conv_tc launches of a step, launch list
mkdir -p gpurun_out /tmp/ncu
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r34_tests.log
grep -E "passed|failed|error" gpurun_out/r34_tests.log | tail -2
timeout 200 python __graft_entry__.py --smoke > gpurun_out/r34_smoke.log 2>&1; tail -1 gpurun_out/r34_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r34_bench_default.json 2> gpurun_out/r34_bench_default.err
cut -c1-400 gpurun_out/r34_bench_default.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r34_bench_reference.json 2> gpurun_out/r34_bench_reference.err
cut -c1-200 gpurun_out/r34_bench_reference.json
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' -s 1377 -c 459 --csv --log-file /tmp/ncu/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r34_ncu_traffic.log 2>&1
python tools/traffic_summary.py /tmp/ncu/traffic.csv > gpurun_out/r34_traffic_conv_tc.json 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'conv_tc_kernel' -s 1377 -c 70 -o /tmp/ncu/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r34_ncu_full.log 2>&1
ncu -i /tmp/ncu/prof.ncu-rep --page raw --csv > gpurun_out/r34_conv_tc_full_raw.csv 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/ncu/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r34_ncu_launches.log 2>&1
python tools/ncu_summary.py launches /tmp/ncu/launches.csv > gpurun_out/r34_launches_summary.txt 2>&1
head -14 gpurun_out/r34_launches_summary.txt
ls -la gpurun_out/ | grep r34
