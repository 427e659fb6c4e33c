#!/bin/bash
# DCN backward bring-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dcn_gpu.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r15_dcn_tests.log
timeout 300 python - > gpurun_out/r15_dcn_bwd_time.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from glare_b200.dcn_backward import dcn_backward
from glare_b200 import ops
for (B, C, H, W) in ((2, 128, 420, 620), (2, 256, 210, 310)):
    g = torch.Generator().manual_seed(1)
    x = torch.randn((B, C, H, W), generator=g).cuda(); off = (torch.randn((B, 72, H, W), generator=g) * 2).cuda()
    m = torch.sigmoid(torch.randn((B, 36, H, W), generator=g)).cuda(); w = (torch.randn((C, C, 3, 3), generator=g) / 34).cuda()
    go = torch.randn((B, C, H, W), generator=g).cuda()
    for _ in range(2): dcn_backward(x, off, m, w, go, 4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): dcn_backward(x, off, m, w, go, 4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    fl = 3 * 2.0 * B * H * W * C * C * 9
    print("DCN backward B=%d C=%d %dx%d: %.2f ms  (%.1f TFLOP/s over the three GEMM-shaped terms)" % (B, C, H, W, ms, fl / ms / 1e9))
PY
tail -6 gpurun_out/r15_dcn_tests.log; cat gpurun_out/r15_dcn_bwd_time.log | tail -3
