"""GPU parity: modulated deformable conv forward through the C ABI vs the golden (torchvision == literal
restatement of the reference CUDA kernel, PIN_REPORT) and the C oracle.  fp32 path tolerance 1e-4 abs/rel."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _dcn(x, off, msk, w, b, dg=4, stride=1, pad=1, dil=1):
    from glare_b200 import ops
    y = ops.modulated_deform_conv(x.cuda(), off.cuda(), msk.cuda(), w.cuda(), None if b is None else b.cuda(), stride, pad, dil, 1, dg)
    torch.cuda.synchronize()
    return y.cpu()


def test_golden(glare_lib):
    g = {k: torch.from_numpy(v) for k, v in load_golden("dcn").items()}
    y = _dcn(g["x"], g["offset"], g["mask"], g["weight"], g["bias"])
    assert torch.allclose(y, g["y"], atol=1e-5, rtol=1e-5), float((y - g["y"]).abs().max())


@pytest.mark.parametrize("cfg", [
    # B, C, Cout, H, W, dg, with_bias
    (1, 4, 4, 1, 1, 4, True), (2, 8, 12, 5, 7, 4, False), (1, 32, 130, 13, 17, 4, True), (1, 64, 64, 23, 19, 2, True),
    (2, 128, 128, 30, 41, 4, True), (1, 256, 256, 16, 24, 4, True)])
def test_against_oracle(glare_lib, cfg):
    from oracle import glare_oracle as O
    B, C, Co, H, W, dg, wb = cfg
    g = torch.Generator().manual_seed(C * 7 + H)
    x = torch.randn((B, C, H, W), generator=g)
    off = torch.randn((B, dg * 18, H, W), generator=g) * 3.0
    off[0, :, 0, 0] = 1000.0
    off[0, 1::2, H - 1, W - 1] = -0.75
    msk = torch.sigmoid(torch.randn((B, dg * 9, H, W), generator=g))
    w = torch.randn((Co, C, 3, 3), generator=g) / (3.0 * C ** 0.5)
    b = torch.randn((Co,), generator=g) if wb else None
    y = _dcn(x, off, msk, w, b, dg)
    y_o = O.modulated_deform_conv(x, off, msk, w, b, dg=dg)
    assert torch.allclose(y, y_o, atol=1e-4, rtol=1e-4), float((y - y_o).abs().max())


def test_stride_dilation_against_oracle(glare_lib):
    from oracle import glare_oracle as O
    g = torch.Generator().manual_seed(9)
    B, C, Co, H, W, dg = 1, 8, 8, 11, 14, 2
    x = torch.randn((B, C, H, W), generator=g)
    Ho, Wo = (H + 2 * 2 - (2 * 2 + 1)) // 2 + 1, (W + 2 * 2 - (2 * 2 + 1)) // 2 + 1
    off = torch.randn((B, dg * 18, Ho, Wo), generator=g) * 1.5
    msk = torch.sigmoid(torch.randn((B, dg * 9, Ho, Wo), generator=g))
    w = torch.randn((Co, C, 3, 3), generator=g) * 0.2
    y = _dcn(x, off, msk, w, None, dg, stride=2, pad=2, dil=2)
    y_o = O.modulated_deform_conv(x, off, msk, w, None, stride=2, padding=2, dilation=2, dg=dg)
    assert torch.allclose(y, y_o, atol=1e-4, rtol=1e-4)


def test_zero_offset_is_masked_conv_full_size(glare_lib):
    """known answer at the AFT scale-1 shape (128ch @ 420x620): offsets 0, mask 0.5 -> 0.5 * conv2d"""
    g = torch.Generator().manual_seed(4)
    x = torch.randn((1, 128, 420, 620), generator=g).cuda()
    w = (torch.randn((128, 128, 3, 3), generator=g) / 34.0).cuda()
    b = torch.randn((128,), generator=g).cuda()
    off = torch.zeros((1, 72, 420, 620), device="cuda")
    msk = torch.full((1, 36, 420, 620), 0.5, device="cuda")
    from glare_b200 import ops
    y = ops.modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 4)
    torch.backends.cudnn.allow_tf32 = False
    ref = 0.5 * F.conv2d(x, w, None, padding=1) + b.view(1, -1, 1, 1)
    assert float((y - ref).abs().max()) < 2e-4


def test_linearity_in_input_full_size(glare_lib):
    """size-independent property at the AFT scale-0 shape (256ch @ 210x310): DCN is linear in x for fixed offsets"""
    from glare_b200 import ops
    g = torch.Generator().manual_seed(8)
    B, C, H, W = 1, 256, 210, 310
    x1 = torch.randn((B, C, H, W), generator=g).cuda()
    x2 = torch.randn((B, C, H, W), generator=g).cuda()
    off = (torch.randn((B, 72, H, W), generator=g) * 4).cuda()
    msk = torch.sigmoid(torch.randn((B, 36, H, W), generator=g)).cuda()
    w = (torch.randn((C, C, 3, 3), generator=g) / 48.0).cuda()
    pk = ops.dcn_pack_weight(w)
    f = lambda x: ops.modulated_deform_conv(x, off, msk, w, None, 1, 1, 1, 1, 4, packed_weight=pk)
    lhs = f(2.0 * x1 - 3.0 * x2)
    rhs = 2.0 * f(x1) - 3.0 * f(x2)
    assert float((lhs - rhs).abs().max()) < 1e-3


def test_bad_shapes_raise(glare_lib):
    from glare_b200 import ops
    x = torch.zeros((1, 8, 4, 4), device="cuda")
    with pytest.raises(ValueError):
        ops.modulated_deform_conv(x, torch.zeros((1, 10, 4, 4), device="cuda"), torch.zeros((1, 36, 4, 4), device="cuda"),
                                  torch.zeros((8, 8, 3, 3), device="cuda"), None, 1, 1, 1, 1, 4)


@pytest.mark.parametrize("mode,tol", [(4, 1e-4), (3, 5e-5), (2, 5e-5), (1, 5e-3), (0, 3e-2)])
@pytest.mark.parametrize("cfg", [(1, 128, 128, 8, 16), (2, 128, 128, 19, 27), (1, 256, 256, 33, 41), (1, 128, 128, 105, 155)])
def test_tensor_core_dcn_pack_against_fp32_kernel(glare_lib, cfg, mode, tol):
    """dcn_tc.cu (sampled A operand + tcgen05) vs the fp32 FMA kernel (itself checked against the oracle above),
    from the RAW conv_offset output: chunk / cat / sigmoid fused (deformableDecoder_arch.py:141-152)"""
    from glare_b200 import ops
    from glare_b200.dense import TcDense
    B, C, Co, H, W = cfg
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn((B, C, H, W), generator=g).cuda()
    raw = torch.randn((B, 108, H, W), generator=g)
    raw[:, :72] *= 3.0
    raw[0, :72, 0, 0] = 500.0                      # far outside
    raw[0, 0:72:2, H - 1, W - 1] = -0.75           # straddles the border
    raw = raw.cuda()
    w = (torch.randn((Co, C, 3, 3), generator=g) / (3.0 * C ** 0.5)).cuda()
    b = torch.randn((Co,), generator=g).cuda()
    o1, o2, m = torch.chunk(raw, 3, dim=1)
    ref = ops.modulated_deform_conv(x, torch.cat((o1, o2), 1), torch.sigmoid(m), w, b, 1, 1, 1, 1, 4)
    d = TcDense(mode)
    y = d.dcn_pack(x.contiguous(memory_format=torch.channels_last), raw.contiguous(memory_format=torch.channels_last), w, b, 4)
    torch.cuda.synchronize()
    assert y is not None
    err = float((y - ref).abs().max())
    assert err < tol * max(1.0, float(ref.abs().max())), (cfg, mode, err)


@pytest.mark.parametrize("cfg", [(1, 32, 32, 7, 9, 4), (2, 64, 64, 13, 17, 4), (1, 128, 128, 24, 31, 4), (1, 64, 32, 5, 6, 2)])
def test_backward_against_cpu_autograd(glare_lib, cfg):
    """grad input / offset / mask / weight / bias of the DCN op (deform_conv.py:155-184) against torchvision's CPU autograd of the
    same operator (== the reference kernels' semantics, PIN_REPORT).  fp32 path; tolerance 2e-4 of each gradient's scale."""
    import torchvision.ops
    from glare_b200.dcn_backward import modulated_deform_conv
    B, C, Co, H, W, dg = cfg
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn((B, C, H, W), generator=g)
    off = torch.randn((B, dg * 18, H, W), generator=g) * 2.0
    off[0, :, 0, 0] = 50.0
    off[0, 0::2, H - 1, 0] = -0.6
    msk = torch.sigmoid(torch.randn((B, dg * 9, H, W), generator=g))
    w = torch.randn((Co, C, 3, 3), generator=g) / (3.0 * C ** 0.5)
    b = torch.randn((Co,), generator=g)
    gout = torch.randn((B, Co, H, W), generator=g)
    ref_in = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
    torchvision.ops.deform_conv2d(ref_in[0], ref_in[1], ref_in[3], ref_in[4], stride=1, padding=1, dilation=1, mask=ref_in[2]).backward(gout)
    dev_in = [t.clone().cuda().requires_grad_(True) for t in (x, off, msk, w, b)]
    y = modulated_deform_conv(dev_in[0], dev_in[1], dev_in[2], dev_in[3], dev_in[4], 1, 1, 1, 1, dg)
    y.backward(gout.cuda())
    torch.cuda.synchronize()
    for name, r, d in zip(("input", "offset", "mask", "weight", "bias"), ref_in, dev_in):
        err = float((d.grad.cpu() - r.grad).abs().max())
        scale = max(1.0, float(r.grad.abs().max()))
        assert err < 2e-4 * scale, (name, cfg, err, scale)


def _case(seed, B=2, C=32, Co=32, H=13, W=17, dg=4):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((B, C, H, W), generator=g)
    off = torch.randn((B, dg * 18, H, W), generator=g) * 2.0
    msk = torch.sigmoid(torch.randn((B, dg * 9, H, W), generator=g))
    w = torch.randn((Co, C, 3, 3), generator=g) / (3.0 * C ** 0.5)
    b = torch.randn((Co,), generator=g)
    gy = torch.randn((B, Co, H, W), generator=g)
    return [t.cuda().contiguous() for t in (x, off, msk, w, b, gy)]


def test_deform_conv_ext_shim_forward_backward(glare_lib):
    """glare_b200/deform_conv_ext.py called the way the reference's ModulatedDeformConvFunction calls its plug-in (deform_conv.py:147-153,
    161-170): caller-allocated output and zeroed gradient tensors, two empty scratch tensors; against torchvision's CPU autograd"""
    import torchvision.ops
    from glare_b200 import deform_conv_ext as ext
    x, off, msk, w, b, gy = _case(21)
    out = x.new_empty((2, 32, 13, 17))
    bufs = [x.new_empty(0), x.new_empty(0)]
    ext.modulated_deform_conv_forward(x, w, b, bufs[0], off, msk, out, bufs[1], 3, 3, 1, 1, 1, 1, 1, 1, 1, 4, True)
    grads = [torch.zeros_like(t) for t in (x, w, b, off, msk)]
    ext.modulated_deform_conv_backward(x, w, b, bufs[0], off, msk, bufs[1], grads[0], grads[1], grads[2], grads[3], grads[4], gy,
                                       3, 3, 1, 1, 1, 1, 1, 1, 1, 4, True)
    assert bufs[0].numel() == 0 and bufs[1].numel() == 0
    cx, coff, cmsk, cw, cb = (t.detach().cpu().requires_grad_(True) for t in (x, off, msk, w, b))
    y_ref = torchvision.ops.deform_conv2d(cx, coff, cw, cb, stride=1, padding=1, dilation=1, mask=cmsk)
    y_ref.backward(gy.cpu())
    assert torch.allclose(out.cpu(), y_ref.detach(), atol=1e-4, rtol=1e-4)
    for got, want, nm in zip(grads, (cx.grad, cw.grad, cb.grad, coff.grad, cmsk.grad), ("input", "weight", "bias", "offset", "mask")):
        sc = max(1.0, float(want.abs().max()))
        assert float((got.cpu() - want).abs().max()) < 2e-3 * sc, nm


def test_against_the_reference_built_extension(glare_lib):
    """oracle/_ref/deform_conv_ext.so = the reference's own CUDA sources compiled unmodified for sm_100 (oracle/build_ref_dcn.py): the
    reference GPU kernel itself as the checker of both glare DCN kernels.  Skipped when the extension was not built."""
    import os
    import sys
    from conftest import ROOT
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "deform_conv_ext.so")):
        pytest.skip("oracle/_ref/deform_conv_ext.so not built (python -m oracle.build_ref_dcn, authoring container)")
    sys.path.insert(0, ref_dir)
    try:
        import deform_conv_ext as ref_ext
    finally:
        sys.path.remove(ref_dir)
    from glare_b200 import ops
    from glare_b200.dense import make_dense
    x, off, msk, w, b, _ = _case(5, B=2, C=128, Co=128, H=30, W=41)
    out = x.new_empty(x.shape)
    ref_ext.modulated_deform_conv_forward(x, w, b, x.new_empty(0), off, msk, out, x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 4, True)
    y_fma = ops.modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, 4)
    assert float((y_fma - out).abs().max()) < 1e-4
    # tensor-core kernel: consumes the raw conv_offset output (offsets in the reference's (g, tap, {dh, dw}) order, mask LOGITS)
    o1, o2 = torch.chunk(off, 2, dim=1)
    raw = torch.cat((o1, o2, torch.logit(msk.clamp(1e-6, 1 - 1e-6))), dim=1)
    y_tc = make_dense("auto").dcn_pack(x, raw, w, b, 4)
    assert float((y_tc - out).abs().max()) < 2e-4 * max(1.0, float(out.abs().max()))
