#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_pipeline_gpu.py tests/test_dcn_gpu.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/r58_bench_$i.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/r58_bench_$i.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['breakdown_ms_per_step'], d['clocks']['sm_mhz'])"
done
