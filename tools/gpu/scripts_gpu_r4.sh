#!/bin/bash
# fourth GPU call: TMA-store epilogue, DCN on tensor cores, full-sample attention bands + register softmax
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=line 2>&1 | tail -40 > gpurun_out/r4_conv_tests.log
timeout 900 python -m pytest tests/test_dcn_gpu.py -m gpu -q --tb=line 2>&1 | tail -40 > gpurun_out/r4_dcn_tests.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_flow_gpu.py tests/test_vq_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r4_pipe_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r4_smoke.log 2>&1
for d in tc-3xtf32 tc-bf16; do
  timeout 600 python bench.py --steps 2 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r4_bench_$d.json 2> gpurun_out/r4_bench_$d.err
done
GLARE_CONV_DIRECT_STORE=1 timeout 600 python bench.py --steps 2 --warmup 3 --dense tc-bf16 --no-cpu-baseline > gpurun_out/r4_bench_tc-bf16_directstore.json 2> gpurun_out/r4_bench_tc-bf16_directstore.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r4_launches_bf16.csv python bench.py --steps 1 --warmup 3 --batch 1 --dense tc-bf16 --no-cpu-baseline > gpurun_out/r4_ncu_bf16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel|dcn_tc_kernel|attn_softmax' -s 150 -c 12 -o gpurun_out/r4_prof_bf16 python bench.py --steps 1 --warmup 3 --batch 1 --dense tc-bf16 --no-cpu-baseline > gpurun_out/r4_ncu_full.log 2>&1
tail -4 gpurun_out/r4_conv_tests.log; tail -4 gpurun_out/r4_dcn_tests.log; tail -4 gpurun_out/r4_pipe_tests.log; tail -2 gpurun_out/r4_smoke.log; cat gpurun_out/r4_bench_*.json | cut -c1-300
