#!/bin/bash
# round 2, call 11: final N=1 bench line (default flags the driver uses: --steps 20 --warmup 5) + reference arm + training kernel-time breakdown
mkdir -p gpurun_out
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r50_bench.json 2> gpurun_out/r50_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r50_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'], d['clocks'])
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r50_bench_reference.json 2> gpurun_out/r50_bench_reference.err; cut -c1-300 gpurun_out/r50_bench_reference.json
timeout 900 python tools/gpu/train_probe.py 3 --profile > gpurun_out/r50_train_probe_profile.txt 2>&1; tail -34 gpurun_out/r50_train_probe_profile.txt
