"""Child process of tests/test_zz_flow_train_gpu.py: the two-ring patch staging of conv_tc for the 256-wide N tiles (default) and the
one-ring kernel (GLARE_CONV_NO_RING2=1) against cuDNN fp32, with CUDA-event timings printed.  The variant is selected by an environment
variable read once per process, so the parent runs this script twice."""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from glare_b200 import ops  # noqa: E402


def main():
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    g = torch.Generator().manual_seed(5)
    for (B, H, W, Ci, Co) in ((2, 105, 155, 256, 256), (1, 210, 310, 512, 256), (4, 105, 155, 512, 512)):
        x = torch.randn((B, Ci, H, W), generator=g).cuda()
        w = (torch.randn((Co, Ci, 3, 3), generator=g) / (3 * Ci ** 0.5)).cuda()
        b = torch.randn((Co,), generator=g).cuda()
        xn = x.permute(0, 2, 3, 1).contiguous()
        w_hi, _ = ops.conv_pack_weight(4, w)
        x_hi, _ = ops.conv_prep_act(4, xn)
        y = ops.conv2d_nhwc_tc(4, x_hi, None, w_hi, None, b, None, B, H, W, Ci, Co, 3).permute(0, 3, 1, 2)
        ref = F.conv2d(x, w, b, padding=1)
        err = float((y - ref).abs().max()) / max(1.0, float(ref.abs().max()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv2d_nhwc_tc(4, x_hi, None, w_hi, None, b, None, B, H, W, Ci, Co, 3)
        e1.record()
        torch.cuda.synchronize()
        good = err < 1.5e-4            # K = 4608 with the tensor core's truncating fp32 accumulator (tests/test_conv_tc_gpu.py uses 1e-4 there)
        ok = ok and good
        print("%-4s RING2=%s conv %dx%dx%d %d->%d: err %.3g of scale, %.3f ms" %
              ("ok" if good else "FAIL", os.environ.get("GLARE_CONV_RING2", "0"), B, H, W, Ci, Co, err, e0.elapsed_time(e1) / 10), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
