#!/bin/bash
# second GPU call: tcgen05 conv + GroupNorm bring-up, pipeline parity per dense backend, benches, ncu of conv/dcn/vq
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --tb=line 2>&1 | tail -80 > gpurun_out/r2_conv_tests.log
timeout 900 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_conv_tc_gpu.py 2>&1 | tail -80 > gpurun_out/r2_tests.log
for d in tc-3xtf32 tc-tf32 tc-bf16; do
  timeout 600 python bench.py --steps 2 --warmup 3 --dense $d --no-cpu-baseline > gpurun_out/r2_bench_$d.json 2> gpurun_out/r2_bench_$d.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -s 40 -c 4 -o gpurun_out/r2_prof_conv python bench.py --steps 1 --warmup 3 --batch 2 --dense tc-3xtf32 --no-cpu-baseline > gpurun_out/r2_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dcn_fwd|vq_argmin|gn_' -s 20 -c 8 -o gpurun_out/r2_prof_misc python bench.py --steps 1 --warmup 3 --batch 2 --dense tc-3xtf32 --no-cpu-baseline > gpurun_out/r2_ncu_misc.log 2>&1
tail -15 gpurun_out/r2_conv_tests.log; tail -8 gpurun_out/r2_tests.log; cat gpurun_out/r2_bench_*.json; tail -3 gpurun_out/r2_bench_*.err
