#!/bin/bash
# round 2, call 3: full-size parity tests, graph-replay test, smoke, bench (graph) with alt_configs, bench eager for the A/B, RING2 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_parity_gpu.py tests/test_api_gpu.py -m gpu -x -q -s > gpurun_out/r42_pytest_parity.log 2>&1; echo "parity rc=$?"; grep -E "420x620|passed|failed|Error|error" gpurun_out/r42_pytest_parity.log | tail -12
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r42_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r42_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r42_bench_graph.json 2> gpurun_out/r42_bench_graph.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/r42_bench_graph.json; tail -5 gpurun_out/r42_bench_graph.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-alt --no-cpu-baseline > gpurun_out/r42_bench_eager.json 2> gpurun_out/r42_bench_eager.err; echo "bench eager rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r42_bench_eager.json').read().strip().splitlines()[-1]); print('eager', d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
GLARE_CONV_RING2=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/r42_bench_ring2.json 2> gpurun_out/r42_bench_ring2.err; echo "bench ring2 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r42_bench_ring2.json').read().strip().splitlines()[-1]); print('ring2', d['value'], d['ms_per_step'], d['breakdown_ms_per_step'], d['clocks'])"
