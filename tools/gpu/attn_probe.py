"""One AttnBlock core (B=1, C=512, 105x155) through the fused-softmax and the three-kernel path: per-kernel CUDA-event times.
Run plain for timings, or under ncu (-k regex:conv_tc_kernel) for the pipe statistics of the two scores GEMMs."""
import sys
import torch
sys.path.insert(0, ".")
from glare_b200.dense import TcDense

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = torch.Generator().manual_seed(0)
q, k, v = (torch.randn((1, 512, 105, 155), generator=g).cuda() for _ in range(3))
for name, fused in (("fused", True), ("unfused", False)):
    d = TcDense(4)
    d.attn_fused = fused
    for _ in range(2):
        d.attention(q, k, v)
    torch.cuda.synchronize()
    d.timers = {}
    for _ in range(reps):
        d.attention(q, k, v)
    torch.cuda.synchronize()
    ev = d.timers
    gem = ev["conv_tc"]
    s_ms = sum(a.elapsed_time(b) for a, b, _ in gem[0::2]) / reps
    pv_ms = sum(a.elapsed_time(b) for a, b, _ in gem[1::2]) / reps
    sm_ms = sum(a.elapsed_time(b) for a, b, _ in ev["attn_softmax"]) / reps
    tot = sum(a.elapsed_time(b) for a, b, _ in ev["attention_total(incl. its conv_tc GEMMs)"]) / reps
    fl = 2.0 * 16275 * 16275 * 512
    print("%-8s scores GEMM %.3f ms (%.0f TFLOP/s alg)  softmax/norm/finish %.3f ms  PV GEMM %.3f ms (%.0f TFLOP/s alg)  block total %.3f ms"
          % (name, s_ms, fl / s_ms / 1e9, sm_ms, pv_ms, fl / pv_ms / 1e9, tot))
