"""DCNv2Pack tail alone (C = 128 at 420x620, the scale-1 WarpBlock) for ncu: python tools/gpu/dcn_probe.py [B]"""
import sys
import torch
sys.path.insert(0, ".")
from glare_b200.dense import TcDense

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator().manual_seed(0)
C, H, W, dg = 128, 420, 620, 4
x = torch.randn((B, C, H, W), generator=g).cuda().contiguous(memory_format=torch.channels_last)
om = (torch.randn((B, 27 * dg, H, W), generator=g) * 1.5).cuda().contiguous(memory_format=torch.channels_last)
w = (torch.randn((C, C, 3, 3), generator=g) / (3 * C ** 0.5)).cuda()
b = torch.randn((C,), generator=g).cuda()
d = TcDense(4)
for _ in range(2):
    y = d.dcn_pack(x, om, w, b, dg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = d.dcn_pack(x, om, w, b, dg)
e1.record()
torch.cuda.synchronize()
print("dcn_tc C=%d %dx%d B=%d: %.3f ms / launch" % (C, H, W, B, e0.elapsed_time(e1) / 5))
