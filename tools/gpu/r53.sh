#!/bin/bash
# round 2, call 14 (8 GPUs): the driver's scaling command at N = 8 (short) -- graph replay per rank, async uint8 gather, alt configs incl. the data-parallel stage-2 step
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r53_bench_8gpu.json 2> gpurun_out/r53_bench_8gpu.err; echo "bench8 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r53_bench_8gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['config']['gather'])
a=d['alt_configs']; print('bf16', a['lolv2_real_bf16_bs64_over_8gpus']['value'], 'train', a['stage2_training_step']['value'], a['stage2_training_step']['ms_per_step'], '1080p', {k:v['value'] for k,v in a['unpaired_1080p_fp32']['per_gpu_batch_sweep'].items()})
PY
tail -3 gpurun_out/r53_bench_8gpu.err
