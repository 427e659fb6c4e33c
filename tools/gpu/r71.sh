# round-2 final sequence after the stage-3 work: full GPU suite, smoke, bench N=1 (alt configs incl. the stage-2 / stage-3 training steps)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r71_pytest_gpu.log 2>&1; tail -3 gpurun_out/r71_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r71_smoke.log 2>&1; tail -2 gpurun_out/r71_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r71_bench.json 2> gpurun_out/r71_bench.err; tail -c 600 gpurun_out/r71_bench.err
python - <<'PY'
import json
for line in open('gpurun_out/r71_bench.json'):
    if line.startswith('{'):
        j = json.loads(line)
        print(j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['traffic'])
        a = j['alt_configs']
        print({k: (v.get('value'), v.get('ms_per_step')) if isinstance(v, dict) else v for k, v in a.items()})
PY
timeout 300 python tools/gpu/train_probe.py 5 --profile > gpurun_out/r71_train_probe_kernel_breakdown.txt 2>&1; grep "encoder forward" gpurun_out/r71_train_probe_kernel_breakdown.txt
GLARE_DENSE=tc-bf16 timeout 300 python tools/gpu/train_probe.py 5 --profile > gpurun_out/r71_train_probe_bf16_kernel_breakdown.txt 2>&1; grep "encoder forward" gpurun_out/r71_train_probe_bf16_kernel_breakdown.txt
timeout 300 python tools/gpu/stage3_probe.py 10 --profile > gpurun_out/r71_stage3_probe.txt 2>&1; grep "stage-3 step" gpurun_out/r71_stage3_probe.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r71_bench_reference.json 2>/dev/null
