#!/bin/bash
# first GPU call of the next round: (1) the training kernels' first hardware run, (2) A/B of the opt-in variants prepared without a GPU
mkdir -p gpurun_out
timeout 600 python tests/flow_train_gpu_check.py > gpurun_out/r35_train_check.log 2>&1; echo "train check rc=$?"; tail -25 gpurun_out/r35_train_check.log
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x -k "conv_against and 4" 2>&1 | tail -3
GLARE_CONV_RING2=1 timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -5
for v in "" "GLARE_CONV_RING2=1"; do
  env $v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r35_bench_${v:-default}.json 2> gpurun_out/r35_bench_${v:-default}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r35_bench_${v:-default}.json").read().strip().splitlines()[-1])
    print("${v:-default}", d["value"], d["ms_per_step"], d["breakdown_ms_per_step"], d["clocks"])
except Exception as e:
    print("${v:-default} failed", e)
PY
done
timeout 900 python tools/gpu/train_probe.py 3 > gpurun_out/r35_train_probe.txt 2>&1; tail -3 gpurun_out/r35_train_probe.txt
GLARE_WGRAD_TC=1 timeout 900 python tools/gpu/train_probe.py 3 > gpurun_out/r35_train_probe_wgrad_tc.txt 2>&1; tail -3 gpurun_out/r35_train_probe_wgrad_tc.txt
