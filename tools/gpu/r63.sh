#!/bin/bash
# round 2, call 26: conv_tc DRAM traffic of one step re-captured for the final kernel sources (hash-tagged), full GPU suite, smoke
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-alt --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' -s 1410 -c 470 --csv --log-file /tmp/ncu/traffic.csv $CMD > gpurun_out/r63_ncu_traffic.log 2>&1; echo "traffic rc=$?"
python tools/traffic_summary.py /tmp/ncu/traffic.csv > gpurun_out/r63_traffic_conv_tc.json 2>&1
python tools/traffic_summary.py /tmp/ncu/traffic.csv --bench-json tcgen05-bf16x3 15 "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 1410 -c 470 $CMD (tools/gpu/r63.sh)" > gpurun_out/conv_tc_traffic.json 2>&1; cat gpurun_out/conv_tc_traffic.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r63_pytest_gpu.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r63_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r63_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r63_smoke.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r63_bench.json 2> gpurun_out/r63_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r63_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['breakdown_ms_per_step'], d['clocks'])
a=d['alt_configs']; print('bf16', a['lolv2_real_bf16_bs64_over_8gpus']['value'], 'train', a['stage2_training_step']['ms_per_step'], '1080p', {k:v['value'] for k,v in a['unpaired_1080p_fp32']['per_gpu_batch_sweep'].items()})
PY
