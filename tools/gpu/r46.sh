#!/bin/bash
# round 2, call 7: RING2 default + fused GN stats: full GPU suite, smoke, bench with alt configs (mirror training step with the batched FlowPlan)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r46_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r46_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r46_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r46_smoke.log
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r46_bench.json 2> gpurun_out/r46_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r46_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown_ms_per_step'], d['clocks'], d['roofline']['frac'])
a=d['alt_configs']; print(a['lolv2_real_bf16_bs64_over_8gpus']['value'], {k:v['value'] for k,v in a['unpaired_1080p_fp32']['per_gpu_batch_sweep'].items()})
print(a['stage2_training_step'])
PY
tail -3 gpurun_out/r46_bench.err
timeout 900 python tools/gpu/train_probe.py 3 > gpurun_out/r46_train_probe.txt 2>&1; tail -2 gpurun_out/r46_train_probe.txt
GLARE_WGRAD_TC=1 timeout 900 python tools/gpu/train_probe.py 3 > gpurun_out/r46_train_probe_tc.txt 2>&1; tail -2 gpurun_out/r46_train_probe_tc.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-alt --no-cpu-baseline --dense tc-bf16 --batch 8 > gpurun_out/r46_bench_bf16_b8.json 2> gpurun_out/r46_bench_bf16_b8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r46_bench_bf16_b8.json').read().strip().splitlines()[-1])
print('bf16 b8', d['value'], d['ms_per_step'], d['breakdown_ms_per_step'])
PY
