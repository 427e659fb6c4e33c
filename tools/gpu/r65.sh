#!/bin/bash
# round 2, final check: the driver's round-end sequence on the final tree (GPU suite, smoke, N=1 bench with alt configs, reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r65_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r65_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r65_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r65_smoke.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r65_bench.json 2> gpurun_out/r65_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r65_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
a=d['alt_configs']; print('bf16', a['lolv2_real_bf16_bs64_over_8gpus']['value'], 'lat', a['latency_600x400_batch1']['value'], 'train', a['stage2_training_step']['ms_per_step'], '1080p', {k:round(v['value'],3) for k,v in a['unpaired_1080p_fp32']['per_gpu_batch_sweep'].items()})
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
