#!/bin/bash
# sixth GPU call: re-verify after the ragged-Cout bias fix + RN tf32 split, API test, smoke, benches, conv_tc DRAM traffic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=line 2>&1 | tail -30 > gpurun_out/r6_conv_tests.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -rP --deselect tests/test_conv_tc_gpu.py 2>&1 | tail -120 > gpurun_out/r6_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r6_smoke.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r6_bench_default.json 2> gpurun_out/r6_bench_default.err
timeout 600 python bench.py --steps 3 --warmup 3 --dense tc-bf16 --no-cpu-baseline > gpurun_out/r6_bench_tc-bf16.json 2> gpurun_out/r6_bench_tc-bf16.err
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel|dcn_tc_kernel' -s 1600 -c 420 --csv --log-file gpurun_out/r6_traffic_3xtf32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu_traffic.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 -s 6800 --csv --log-file gpurun_out/r6_launches_3xtf32.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r6_ncu_launches.log 2>&1
tail -4 gpurun_out/r6_conv_tests.log; grep -E "passed|failed" gpurun_out/r6_tests.log | tail -3; tail -2 gpurun_out/r6_smoke.log; cat gpurun_out/r6_bench_*.json | cut -c1-250
