#!/bin/bash
# first GPU bring-up: parity tests, smoke, bench, ncu launch list + one full capture of each of our kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1_env.txt 2>&1
nproc >> gpurun_out/r1_env.txt; free -g | head -2 >> gpurun_out/r1_env.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r1_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r1_smoke.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'vq_|flow_|dcn_' -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dcn_fwd|flow_tail|vq_argmin' -s 12 -c 6 -o gpurun_out/r1_prof python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
tail -5 gpurun_out/r1_tests.log; cat gpurun_out/r1_smoke.log | tail -3; cat gpurun_out/r1_bench.json
