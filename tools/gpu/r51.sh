#!/bin/bash
# round 2, call 12: fused softmax for the bf16 mode (config 3): tests + bf16 bench breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_fullsize_parity_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/r51_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r51_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-alt --no-cpu-baseline --dense tc-bf16 --batch 8 > gpurun_out/r51_bench_bf16_b8.json 2> gpurun_out/r51_bench_bf16_b8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r51_bench_bf16_b8.json').read().strip().splitlines()[-1])
print('bf16 b8', d['value'], d['ms_per_step'], d['breakdown_ms_per_step'], d['config']['library_fallbacks_per_run'])
PY
tail -3 gpurun_out/r51_bench_bf16_b8.err
timeout 900 python -m pytest tests/test_zz_flow_train_gpu.py tests/test_modules_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/gpu/train_probe.py 3 --profile > gpurun_out/r51_train_probe_profile.txt 2>&1; grep -A12 "stage-2 step" gpurun_out/r51_train_probe_profile.txt | head -16
