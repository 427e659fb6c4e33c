#!/bin/bash
# round 2, call 6: deterministic fused GN statistics, DCN chunk-major K order, batched FlowPlan: tests + A/B benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_dcn_gpu.py -m gpu -x -q -k "groupnorm or dcn or golden or oracle or shim or reference" > gpurun_out/r45_pytest_a.log 2>&1; echo "gn/dcn rc=$?"; tail -4 gpurun_out/r45_pytest_a.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r45_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r45_pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 8 --warmup 3 --no-alt --no-cpu-baseline > gpurun_out/r45_bench_$name.json 2> gpurun_out/r45_bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r45_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"],3), round(d["ms_per_step"],1), d["breakdown_ms_per_step"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name failed", e)
PY
}
run default A=1
run nofusegn GLARE_NO_FUSE_GN_STATS=1
run dcnstages3 GLARE_DCN_STAGES=3
run dcnstages2 GLARE_DCN_STAGES=2
run ring2 GLARE_CONV_RING2=1
timeout 600 python tools/gpu/dcn_ref_compare.py 15 > gpurun_out/r45_dcn_ref_compare.txt 2>&1; grep "dcn_tc_kernel" gpurun_out/r45_dcn_ref_compare.txt
timeout 900 python tools/gpu/train_probe.py 3 > gpurun_out/r45_train_probe.txt 2>&1; tail -2 gpurun_out/r45_train_probe.txt
