#!/bin/bash
# A/B: GroupNorm statistics in the conv epilogue, re-measured with the pipelined epilogue
mkdir -p gpurun_out
GLARE_FUSE_GN_STATS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r27_bench_gnfused.json 2> gpurun_out/r27_bench_gnfused.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r27_bench_gnfused.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["breakdown_ms_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r27_bench_gnfused.err").read()[-2000:])
PY
