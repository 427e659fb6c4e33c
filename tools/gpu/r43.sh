#!/bin/bash
# round 2, call 4: sampled-max softmax reference tests, module mirrors on the GPU, full GPU suite, reference DCN extension beside ours, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r43_pytest_attention.log 2>&1; echo "attention rc=$?"; tail -4 gpurun_out/r43_pytest_attention.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_api_gpu.py -m gpu -x -q > gpurun_out/r43_pytest_modules.log 2>&1; echo "modules rc=$?"; tail -6 gpurun_out/r43_pytest_modules.log
timeout 600 python tools/gpu/dcn_ref_compare.py 15 > gpurun_out/r43_dcn_ref_compare.txt 2>&1; echo "dcn ref rc=$?"; tail -12 gpurun_out/r43_dcn_ref_compare.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r43_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r43_pytest_gpu.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r43_bench.json 2> gpurun_out/r43_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r43_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown_ms_per_step'], d['clocks'])
a=d['alt_configs']; print(a['lolv2_real_bf16_bs64_over_8gpus']['value'], {k:v['value'] for k,v in a['unpaired_1080p_fp32']['per_gpu_batch_sweep'].items()})
PY
tail -3 gpurun_out/r43_bench.err
