"""The reference's OWN CUDA extension (deform_conv_ext, built unmodified for sm_100 into oracle/_ref by oracle/build_ref_dcn.py) beside glare's
DCNv2 kernels on the two AFT scales of the bench workload (B = 15; C = 256 @ 210x310 and C = 128 @ 420x620; deformable_groups 4, fp32 as the
reference forces for the DCN inputs, deformableDecoder_arch.py:143): parity of the three implementations and device time of each.
SURVEY 8d "reference GPU as shipped" column for row a-3.   python tools/gpu/dcn_ref_compare.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import deform_conv_ext  # noqa: E402  (the reference's pybind11 module)

from glare_b200 import ops  # noqa: E402
from glare_b200.dense import make_dense  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 15
dev = torch.device("cuda:0")
dense = make_dense("auto")


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for C, H, W in ((256, 210, 310), (128, 420, 620)):
    g = torch.Generator().manual_seed(C)
    x = torch.randn((B, C, H, W), generator=g).to(dev)
    raw = torch.randn((B, 108, H, W), generator=g).to(dev) * 1.5          # conv_offset output: 72 offsets + 36 mask logits
    wgt = (torch.randn((C, C, 3, 3), generator=g) / (3 * C ** 0.5)).to(dev)
    bias = torch.randn((C,), generator=g).to(dev)
    o1, o2, m = torch.chunk(raw, 3, dim=1)
    offset, mask = torch.cat((o1, o2), dim=1).contiguous(), torch.sigmoid(m).contiguous()

    def ref():
        out = x.new_empty((B, C, H, W))
        bufs = [x.new_empty(0), x.new_empty(0)]
        deform_conv_ext.modulated_deform_conv_forward(x, wgt, bias, bufs[0], offset, mask, out, bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, 4, True)
        return out

    def ref_with_glue():          # what DCNv2Pack.forward runs after conv_offset: chunk / cat / sigmoid + the op
        a, b, mm = torch.chunk(raw, 3, dim=1)
        off, msk = torch.cat((a, b), dim=1), torch.sigmoid(mm)
        out = x.new_empty((B, C, H, W))
        deform_conv_ext.modulated_deform_conv_forward(x, wgt, bias, x.new_empty(0), off, msk, out, x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 4, True)
        return out

    packed = ops.dcn_pack_weight(wgt)
    x_cl, raw_cl = x.contiguous(memory_format=torch.channels_last), raw.contiguous(memory_format=torch.channels_last)
    fma = lambda: ops.modulated_deform_conv(x, offset, mask, wgt, bias, 1, 1, 1, 1, 4, packed_weight=packed)      # noqa: E731
    tc = lambda: dense.dcn_pack(x_cl, raw_cl, wgt, bias, 4)                                                         # noqa: E731
    op = lambda: ops.modulated_deform_conv(x, offset, mask, wgt, bias, 1, 1, 1, 1, 4)        # noqa: E731  (operator-level entry: NCHW tensors of the reference op)
    y_ref, y_fma, y_tc = ref(), fma(), tc()
    y_op, t_op = op(), timed(op)
    sc = float(y_ref.abs().max())
    t_ref, t_glue, t_fma, t_tc = timed(ref), timed(ref_with_glue), timed(fma), timed(tc)
    fl = 2.0 * B * H * W * C * C * 9 / 1e12
    print("DCNv2 C=%d %dx%d B=%d (%.2f algorithmic TFLOP): maxdiff vs reference extension: fp32-FMA kernel %.3g, tensor-core kernel %.3g (|y| max %.3g)"
          % (C, H, W, B, fl, float((y_fma - y_ref).abs().max()), float((y_tc - y_ref).abs().max()), sc))
    print("    reference deform_conv_ext (sm_100 build)          %8.2f ms   %6.1f TFLOP/s   (+ chunk/cat/sigmoid glue: %.2f ms)" % (t_ref, fl / t_ref * 1e3, t_glue))
    print("    glare dcn_fwd_kernel (fp32 FMA, general operator)  %8.2f ms   %6.1f TFLOP/s   x%.1f" % (t_fma, fl / t_fma * 1e3, t_ref / t_fma))
    print("    glare dcn_tc_kernel (tcgen05, raw conv_offset in)  %8.2f ms   %6.1f TFLOP/s   x%.1f (x%.1f incl. the reference's glue)" %
          (t_tc, fl / t_tc * 1e3, t_ref / t_tc, t_glue / t_tc))
    print("    glare modulated_deform_conv (reference op signature, NCHW in/out, same kernel + layout passes) %8.2f ms   x%.1f   maxdiff %.3g" %
          (t_op, t_ref / t_op, float((y_op - y_ref).abs().max())))
    del x, raw, wgt, y_ref, y_fma, y_tc, y_op
    torch.cuda.empty_cache()
