"""GPU parity: tcgen05 implicit-GEMM convolution and the GroupNorm/swish operand kernels through the C ABI,
against cuDNN fp32 (TF32 off) on the same inputs.  Tolerances: 3xTF32 2e-5 relative to the output scale
(fp32-grade), TF32 3e-3, bf16 exact-product check against bf16-rounded operands 2e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [
    # B, H, W, Cin, Cout, ks
    (1, 8, 16, 64, 64, 3),        # exactly one tile
    (1, 8, 16, 64, 64, 1),
    (2, 13, 21, 128, 128, 3),     # ragged tiles, halo across tile borders
    (1, 24, 40, 256, 256, 3),     # BN = 256
    (1, 17, 33, 512, 512, 3),
    (1, 9, 20, 512, 256, 3),
    (1, 26, 39, 256, 108, 3),     # conv_offset: Cout not a multiple of the N tile
    (1, 12, 20, 64, 3072, 3),     # flow pre-pass
    (2, 15, 19, 512, 512, 1),     # attention q/k/v/proj
    (1, 5, 7, 128, 256, 1),       # nin_shortcut
    (1, 105, 155, 128, 128, 3),   # many tiles per CTA: exercises the persistent loop, both TMEM accumulators, stage wrap
]


def _ref(x, w, b, res, ks):
    torch.backends.cudnn.allow_tf32 = False
    y = F.conv2d(x, w, b, padding=ks // 2)
    return y if res is None else y + res


@pytest.mark.parametrize("mode", [4, 3, 2, 1, 0])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_against_cudnn_fp32(glare_lib, shape, mode):
    from glare_b200 import ops
    B, H, W, Ci, Co, ks = shape
    g = torch.Generator().manual_seed(Ci + Co + H)
    x = torch.randn((B, Ci, H, W), generator=g).cuda()
    w = (torch.randn((Co, Ci, ks, ks), generator=g) / (ks * Ci ** 0.5)).cuda()
    b = torch.randn((Co,), generator=g).cuda()
    res = torch.randn((B, Co, H, W), generator=g).cuda()
    xn = x.permute(0, 2, 3, 1).contiguous()
    rn = res.permute(0, 2, 3, 1).contiguous()
    w_hi, w_lo = ops.conv_pack_weight(mode, w)
    x_hi, x_lo = ops.conv_prep_act(mode, xn)
    y = ops.conv2d_nhwc_tc(mode, x_hi, x_lo, w_hi, w_lo, b, rn, B, H, W, Ci, Co, ks).permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    if mode == 0:
        ref = _ref(x.bfloat16().float(), w.bfloat16().float(), b, res, ks)
        tol = 2e-4
    else:
        ref = _ref(x, w, b, res, ks)
        tol = {1: 3e-3, 2: 2e-5, 3: 2e-5, 4: 6e-5}[mode]            # mode 4: 16-bit-significand products
    err = float((y - ref).abs().max())
    assert err < tol * max(1.0, float(ref.abs().max())), (shape, mode, err)
    # no bias / no residual path
    y2 = ops.conv2d_nhwc_tc(mode, x_hi, x_lo, w_hi, w_lo, None, None, B, H, W, Ci, Co, ks).permute(0, 3, 1, 2)
    ref2 = ref - b.view(1, -1, 1, 1) - res
    assert float((y2 - ref2).abs().max()) < tol * max(1.0, float(ref.abs().max()))


def test_unsupported_shapes_are_refused(glare_lib):
    from glare_b200 import ops
    x = torch.zeros((1, 4, 4, 3), device="cuda")
    w = torch.zeros((64, 9, 3), device="cuda")
    with pytest.raises(RuntimeError):
        ops.conv2d_nhwc_tc(1, x, None, w, None, None, None, 1, 4, 4, 3, 64, 3)


@pytest.mark.parametrize("C,H,W,B", [(128, 9, 14, 2), (256, 31, 17, 1), (512, 105, 155, 1), (128, 420, 620, 1)])
@pytest.mark.parametrize("swish", [True, False])
def test_groupnorm_swish_operands(glare_lib, C, H, W, B, swish):
    from glare_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn((B, C, H, W), generator=g) * 2.0 + 3.0).cuda()          # large mean: stresses the variance computation
    gamma = (1 + 0.1 * torch.randn((C,), generator=g)).cuda()
    beta = (0.1 * torch.randn((C,), generator=g)).cuda()
    ref = F.group_norm(x.double(), 32, gamma.double(), beta.double(), eps=1e-6)
    if swish:
        ref = ref * torch.sigmoid(ref)
    xn = x.permute(0, 2, 3, 1).contiguous()
    stats = ops.gn_stats(xn, B, H * W, C)
    for mode, tol in ((4, 2e-5), (3, 2e-6), (2, 2e-6), (1, 2e-6), (0, 1e-2)):
        hi, lo = ops.gn_apply(mode, xn, stats, gamma, beta, swish, B, H * W, C)
        if mode == 4:                      # one interleaved bf16 tensor: [32 x a1 | 32 x a2] per 32-channel chunk, y = a1 + a2
            xx = hi.view(-1, 2, 32).double()
            val = (xx[:, 0] + xx[:, 1]).reshape(B, H, W, C)
        elif mode == 3:                      # interleaved bf16 x tensor: [32 x bf16(y) | 32 x bf16(y - hi)] per 32-channel chunk
            xx = lo.view(-1, 2, 32).double()
            assert float((xx[:, 0].reshape(hi.shape) - hi.double()).abs().max()) < 1e-2 * max(1.0, float(ref.abs().max()))
            val = hi.double() + xx[:, 1].reshape(hi.shape)
        else:
            val = hi.double() if lo is None else hi.double() + lo.double()
        err = float((val.permute(0, 3, 1, 2) - ref).abs().max())
        assert err < tol * max(1.0, float(ref.abs().max())), (mode, err)
        if mode in (2, 3):
            assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0      # hi is an exact tf32 value


@pytest.mark.parametrize("mode,tol", [(4, 6e-5), (3, 2e-5), (2, 2e-5), (1, 3e-3), (0, 2e-2)])
@pytest.mark.parametrize("shape", [(1, 512, 9, 14), (2, 512, 24, 21), (1, 512, 105, 155)])
def test_attention_gemm_path(glare_lib, shape, mode, tol):
    """AttnBlock core (encoder_decoder.py:176-187) through the tcgen05 GEMMs + fused softmax kernel vs an fp64 evaluation"""
    from glare_b200.dense import TcDense
    B, C, h, w = shape
    g = torch.Generator().manual_seed(h * w)
    q = torch.randn((B, C, h, w), generator=g).cuda()
    k = torch.randn((B, C, h, w), generator=g).cuda()
    v = torch.randn((B, C, h, w), generator=g).cuda()
    d = TcDense(mode)
    out = d.attention(q, k, v)
    torch.cuda.synchronize()
    assert d.attention_verified()
    ref = _attention_fp64(q, k, v)
    err = float((out.reshape(B, C, h * w).double() - ref).abs().max())
    assert err < tol * max(1.0, float(ref.abs().max())), (shape, mode, err)


def _attention_fp64(q, k, v):
    B, C, h, w = q.shape
    N = h * w
    ref = torch.empty((B, C, N), dtype=torch.float64, device="cuda")
    for b in range(B):
        qq, kk, vv = q[b].reshape(C, N).double(), k[b].reshape(C, N).double(), v[b].reshape(C, N).double()
        for i0 in range(0, N, 4096):                         # query chunks: keep the fp64 score matrix small
            s = torch.softmax(qq[:, i0:i0 + 4096].t() @ kk * (int(C) ** -0.5), dim=1)
            ref[b, :, i0:i0 + 4096] = vv @ s.t()
    return ref


@pytest.mark.parametrize("shape", [(1, 512, 9, 14), (2, 512, 24, 21), (1, 512, 105, 155), (1, 128, 40, 33)])
def test_attention_fused_softmax_matches_exact_path(glare_lib, shape):
    """mode 4: exp in the scores-GEMM epilogue + 1/rowsum in the P V epilogue vs the three-kernel path (S, softmax, P V) and fp64;
    large logits (|q||k| C^-0.5 up to ~60 with peaked rows) stay inside the safe window"""
    from glare_b200.dense import TcDense
    B, C, h, w = shape
    g = torch.Generator().manual_seed(7 + h)
    q = (torch.randn((B, C, h, w), generator=g) * 2.0).cuda()
    k = (torch.randn((B, C, h, w), generator=g) * 1.5).cuda()
    v = torch.randn((B, C, h, w), generator=g).cuda()
    fused, exact = TcDense(4), TcDense(4)
    exact.attn_fused = False
    assert fused.attn_fused
    o1 = fused.attention(q, k, v)
    o2 = exact.attention(q, k, v)
    torch.cuda.synchronize()
    assert fused.attention_verified() and fused.attn_fused
    ref = _attention_fp64(q, k, v)
    sc = max(1.0, float(ref.abs().max()))
    e1 = float((o1.reshape(B, C, h * w).double() - ref).abs().max())
    e2 = float((o2.reshape(B, C, h * w).double() - ref).abs().max())
    assert e1 < 6e-5 * sc and e2 < 6e-5 * sc, (shape, e1, e2)
    # bitwise repeatable (fixed summation order of the partial row sums)
    o3 = fused.attention(q, k, v)
    assert torch.equal(o1, o3)


@pytest.mark.parametrize("shape", [(2, 13, 21, 128, 128, 3), (1, 24, 40, 512, 512, 1), (1, 105, 155, 256, 128, 3)])
def test_conv_operand_epilogue(glare_lib, shape):
    """conv whose epilogue writes the next GEMM's bf16x3 operand (+ per-row sums of squares): bitwise the operand that the conversion pass
    makes from the fp32 conv output"""
    from glare_b200 import ops
    B, H, W, Ci, Co, ks = shape
    g = torch.Generator().manual_seed(Ci + Co + H)
    x = torch.randn((B, H, W, Ci), generator=g).cuda()
    w = (torch.randn((Co, Ci, ks, ks), generator=g) / (ks * Ci ** 0.5)).cuda()
    b = torch.randn((Co,), generator=g).cuda()
    w_hi, _ = ops.conv_pack_weight(4, w)
    x_hi, _ = ops.conv_prep_act(4, x)
    y = ops.conv2d_nhwc_tc(4, x_hi, None, w_hi, None, b, None, B, H, W, Ci, Co, ks)
    y_op, _ = ops.conv_prep_act(4, y)
    packed, (part, nb) = ops.conv2d_nhwc_tc_pack(4, x_hi, w_hi, b, B, H, W, Ci, Co, ks, row_sq=True)
    torch.cuda.synchronize()
    assert torch.equal(packed.view(torch.int16), y_op.view(torch.int16))
    sq = part[:nb].sum(0).double()
    ref = (y.double() ** 2).sum(-1).reshape(-1)
    assert float((sq - ref).abs().max()) < 1e-5 * float(ref.max())
    norm = torch.empty((B * H * W,), device="cuda")
    mx = torch.zeros((B,), device="cuda", dtype=torch.int32)
    ops.attn_row_norm_finish(part, nb, B * H * W, H * W, norm_out=norm, max_bits=mx)
    assert float((norm.double() - ref.sqrt()).abs().max()) < 2e-5 * float(ref.sqrt().max())
    ref_max = ref.sqrt().reshape(B, -1).max(1).values
    assert float((mx.view(torch.float32).double() - ref_max).abs().max()) < 2e-5 * float(ref_max.max())


def test_attention_block_operand_route(glare_lib):
    """q / k written as operands by their convs and the attention output written as proj_out's operand (engine.attn_block route) against the
    route through fp32 tensors"""
    from glare_b200.dense import TcDense, Operand
    B, C, h, w = 2, 512, 24, 21
    g = torch.Generator().manual_seed(9)
    x = torch.randn((B, C, h, w), generator=g).cuda()
    wq, wk = ((torch.randn((C, C, 1, 1), generator=g) / C ** 0.5).cuda() for _ in range(2))
    bq, bk = (torch.randn((C,), generator=g).cuda() for _ in range(2))
    v = torch.randn((B, C, h, w), generator=g).cuda()
    d = TcDense(4)
    q_op, k_op = d.conv2d_operand(x, wq, bq, row_sq=True), d.conv2d_operand(x, wk, bk, row_sq=True)
    assert isinstance(q_op, Operand) and q_op.row_sq is not None
    o_op = d.attention(q_op, k_op, v, as_operand=True)
    assert isinstance(o_op, Operand)
    o_ref = d.attention(d.conv2d(x, wq, bq, padding=0), d.conv2d(x, wk, bk, padding=0), v)
    assert d.attention_verified()
    ref = _attention_fp64(torch.nn.functional.conv2d(x.double(), wq.double(), bq.double()).float(),
                          torch.nn.functional.conv2d(x.double(), wk.double(), bk.double()).float(), v)
    sc = max(1.0, float(ref.abs().max()))
    assert float((o_op.dense().reshape(B, C, h * w).double() - ref).abs().max()) < 6e-5 * sc
    assert float((o_op.dense() - o_ref).abs().max()) < 2e-5 * sc


@pytest.mark.parametrize("fused", [True, False])
def test_attention_query_bands(glare_lib, fused):
    """score matrices larger than the budget (1080p: 131 648 tokens) are processed in bands of query rows; forced here on a small shape"""
    from glare_b200.dense import TcDense
    B, C, h, w = 2, 512, 24, 21
    g = torch.Generator().manual_seed(5)
    q, k, v = (torch.randn((B, C, h, w), generator=g).cuda() for _ in range(3))
    d = TcDense(4)
    d.attn_fused = fused
    whole = d.attention(q, k, v).clone()
    d.attn_s_budget = 8 * w * 512 * 4                  # 8 image rows of queries per pass -> 3 bands
    banded = d.attention(q, k, v)
    assert d.attention_verified()
    ref = _attention_fp64(q, k, v)
    err = float((banded.reshape(B, C, h * w).double() - ref).abs().max())
    assert err < 6e-5 * max(1.0, float(ref.abs().max())), err
    assert float((banded - whole).abs().max()) < 1e-5


@pytest.mark.parametrize("shape", [(1, 512, 9, 14), (2, 512, 24, 21), (1, 512, 105, 155), (1, 128, 40, 33)])
def test_attention_fused_softmax_bf16_operands(glare_lib, shape):
    """mode 0 (BASELINE config 3): the exp epilogue writes P~ as the single-piece bf16 operand (two 32-key chunks per staged row, ragged
    tails zero-filled) -- against the exact three-kernel bf16 path and fp64, at bf16 tolerances"""
    from glare_b200.dense import TcDense
    B, C, h, w = shape
    g = torch.Generator().manual_seed(3 + h)
    q = (torch.randn((B, C, h, w), generator=g) * 1.5).cuda()
    k = (torch.randn((B, C, h, w), generator=g) * 1.5).cuda()
    v = torch.randn((B, C, h, w), generator=g).cuda()
    fused, exact = TcDense(0), TcDense(0)
    exact.attn_fused = False
    assert fused.attn_fused
    o1, o2 = fused.attention(q, k, v), exact.attention(q, k, v)
    torch.cuda.synchronize()
    assert fused.attention_verified() and fused.attn_fused
    ref = _attention_fp64(q, k, v)
    sc = max(1.0, float(ref.abs().max()))
    e1 = float((o1.reshape(B, C, h * w).double() - ref).abs().max())
    e2 = float((o2.reshape(B, C, h * w).double() - ref).abs().max())
    assert e1 < 3e-2 * sc and e2 < 3e-2 * sc and e1 < 2.0 * e2 + 1e-3, (shape, e1, e2)
    assert torch.equal(o1, fused.attention(q, k, v))


@pytest.mark.parametrize("ref", ["sampled", "cauchy"])
def test_attention_fused_softmax_window_flag_and_fallback(glare_lib, ref):
    """rows outside the fused path's window raise the device flag; attention_verified() then switches the backend to the exact path, which is
    what the engine re-runs with.  "cauchy": row maximum > ~115 below the |q||k| bound (huge norms along orthogonal directions).  "sampled":
    the true row maximum > ~128 above the maximum over the sampled keys (one enormous key that the stride misses)."""
    from glare_b200.dense import TcDense
    B, C, h, w = 1, 512, 12, 16
    g = torch.Generator().manual_seed(3)
    q = torch.randn((B, C, h, w), generator=g) * 0.05
    k = torch.randn((B, C, h, w), generator=g) * 0.05
    if ref == "cauchy":
        q[:, 0] = 400.0                               # bound ~ 400 * 400 / 22.6, logits ~ 0
        k[:, 0] = 0.0
        k[:, 1] = 400.0
        q[:, 1] = 0.0
    else:
        q[:, 0] = 60.0                                # key (0, 1) is not among the 128 sampled of 192 (linspace picks 0, 2, 3, 5, ...)
        k[:, 0] = 0.0
        k[0, 0, 0, 1] = 60.0 * 22.627417 * 0.1        # its logit: 60 * 135.8 / 22.6 = 360 above every sampled key's
    v = torch.randn((B, C, h, w), generator=g)
    q, k, v = q.cuda(), k.cuda(), v.cuda()
    d = TcDense(4)
    d.attn_ref = ref
    if ref == "sampled":
        sel = torch.linspace(0, h * w - 1, 128).round().long()
        assert 1 not in sel.tolist()
    d.attention(q, k, v)
    assert not d.attention_verified()
    assert not d.attn_fused and d.fallbacks
    out = d.attention(q, k, v)
    assert d.attention_verified()
    ref64 = _attention_fp64(q, k, v)
    assert float((out.reshape(B, C, h * w).double() - ref64).abs().max()) < 6e-5 * max(1.0, float(ref64.abs().max()))


@pytest.mark.parametrize("shape,sq,sk", [((1, 512, 24, 21), 3.0, 3.0), ((2, 512, 40, 33), 6.0, 2.0), ((1, 512, 105, 155), 4.0, 4.0)])
def test_attention_fused_softmax_trained_net_logit_ranges(glare_lib, shape, sq, sk):
    """logit ranges of a TRAINED net -- large |q|, |k| in random directions, so that the Cauchy-Schwarz bound
    |q||k| C^-1/2 (~ sq * sk * 22.6) sits far more than 115 above the actual row maximum (~ 4.5 * sq * sk) with peaked rows.  The sampled
    reference keeps these on the fused path (no flag, no fallback) and matches the exact softmax; the round-1 reference trips."""
    from glare_b200.dense import TcDense
    B, C, h, w = shape
    g = torch.Generator().manual_seed(11 + h)
    q = (torch.randn((B, C, h, w), generator=g) * sq).cuda()
    k = (torch.randn((B, C, h, w), generator=g) * sk).cuda()
    v = torch.randn((B, C, h, w), generator=g).cuda()
    N = h * w
    s0 = (q[0].reshape(C, N).t()[:256].double() @ k[0].reshape(C, N).double()) * C ** -0.5
    bound = float(q[0].reshape(C, N).norm(dim=0).max() * k[0].reshape(C, N).norm(dim=0).max()) * C ** -0.5
    assert bound - float(s0.max(dim=1).values.min()) > 130          # the round-1 window is exceeded ...
    d = TcDense(4)
    assert d.attn_ref == "sampled"
    out = d.attention(q, k, v)
    assert d.attention_verified() and d.attn_fused and not d.fallbacks
    ref64 = _attention_fp64(q, k, v)
    sc = max(1.0, float(ref64.abs().max()))
    assert float((out.reshape(B, C, N).double() - ref64).abs().max()) < 1e-4 * sc
    old = TcDense(4)
    old.attn_ref = "cauchy"
    old.attention(q, k, v)
    assert not old.attention_verified()                              # ... and the round-1 reference falls back


@pytest.mark.parametrize("mode,tol", [(4, 1e-4), (3, 5e-5), (2, 5e-5), (1, 3e-3), (0, 2e-4)])   # K = 4608 with a truncating fp32 accumulator
def test_dense_backend_covers_small_channel_and_stride2_convs(glare_lib, mode, tol):
    """TcDense: 3-channel convs (channels zero-padded to the K chunk, 3-channel heads through a padded pixel stride) and
    Downsample (encoder_decoder.py:68-72: pad (0,1,0,1) + stride 2) on the tcgen05 kernel, vs cuDNN fp32"""
    from glare_b200.dense import TcDense
    torch.backends.cudnn.allow_tf32 = False
    d = TcDense(mode)
    g = torch.Generator().manual_seed(17)
    for (Ci, Co, H, W, ks) in [(3, 128, 21, 30, 3), (512, 3, 13, 19, 3), (3, 3, 9, 11, 1), (3, 64, 16, 16, 3)]:
        x = torch.randn((2, Ci, H, W), generator=g).cuda()
        w = (torch.randn((Co, Ci, ks, ks), generator=g) / (ks * Ci ** 0.5)).cuda()
        b = torch.randn((Co,), generator=g).cuda()
        y = d.conv2d(x, w, b, stride=1, padding=ks // 2)
        if mode == 0:
            x, w = x.bfloat16().float(), w.bfloat16().float()
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=ks // 2)       # fp64: cuDNN's fp32 pick for tiny Cout is ~1e-4
        assert float((y - ref).abs().max()) < tol * max(1.0, float(ref.abs().max())), (Ci, Co, ks)
    for (C, Co, H, W) in [(128, 128, 20, 31), (256, 256, 21, 30), (128, 128, 420, 620)]:
        x = torch.randn((1, C, H, W), generator=g).cuda()
        w = (torch.randn((Co, C, 3, 3), generator=g) / (3 * C ** 0.5)).cuda()
        b = torch.randn((Co,), generator=g).cuda()
        y = d.downsample_conv(x, w, b)
        if mode == 0:
            x, w = x.bfloat16().float(), w.bfloat16().float()
        ref = F.conv2d(F.pad(x, (0, 1, 0, 1)).double(), w.double(), b.double(), stride=2)
        assert y.shape == ref.shape
        assert float((y - ref).abs().max()) < tol * max(1.0, float(ref.abs().max())), (C, H, W)
    assert not d.fallbacks


@pytest.mark.parametrize("mode,tol", [(4, 6e-5), (3, 2e-5), (2, 2e-5), (1, 3e-3), (0, 2e-2)])
def test_upsample_conv_subpixel_phases(glare_lib, mode, tol):
    """Upsample.forward (encoder_decoder.py:49-53) through four 2x2 phase convolutions on the low-resolution input vs
    interpolate(nearest, x2) + conv2d in fp64"""
    from glare_b200.dense import TcDense
    d = TcDense(mode)
    g = torch.Generator().manual_seed(23)
    for (C, Co, H, W) in [(64, 64, 8, 16), (256, 256, 13, 21), (512, 512, 105, 155)]:
        x = torch.randn((1, C, H, W), generator=g).cuda()
        w = (torch.randn((Co, C, 3, 3), generator=g) / (3 * C ** 0.5)).cuda()
        b = torch.randn((Co,), generator=g).cuda()
        y = d.upsample_conv(x, w, b)
        ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest").double(), w.double(), b.double(), padding=1)
        assert y is not None and y.shape == ref.shape
        err = float((y - ref).abs().max())
        assert err < tol * max(1.0, float(ref.abs().max())), (C, H, W, err)


@pytest.mark.parametrize("mode,C,H,W", [(4, 128, 22, 37), (4, 256, 33, 18), (4, 512, 9, 21), (3, 128, 22, 37), (0, 128, 22, 37)])
def test_groupnorm_statistics_fused_into_conv_epilogue(glare_lib, mode, C, H, W):
    """conv -> GroupNorm+swish with the statistics taken from the conv epilogue (per-tile partials + fp64 finish) == the same with the
    separate statistics kernel, for the stride-1 (3x3 with residual, halo and standard tiles; 1x1) and Downsample paths; ragged tile edges"""
    from glare_b200.dense import TcDense
    g = torch.Generator().manual_seed(31)
    B = 2
    x = torch.randn((B, C, H, W), generator=g).cuda()
    res = torch.randn((B, C, H, W), generator=g).cuda()
    w = (torch.randn((C, C, 3, 3), generator=g) / (3 * C ** 0.5)).cuda()
    b = torch.randn((C,), generator=g).cuda()
    gamma = (1 + 0.1 * torch.randn((C,), generator=g)).cuda()
    beta = (0.1 * torch.randn((C,), generator=g)).cuda()
    d1, d2 = TcDense(mode), TcDense(mode)
    d1.fuse_gn_stats, d2.fuse_gn_stats = True, False
    w1 = (torch.randn((C, C, 1, 1), generator=g) / C ** 0.5).cuda()
    for fn in (lambda d: d.conv2d(x, w, b, residual=res), lambda d: d.conv2d(x, w1, b, padding=0, residual=res), lambda d: d.downsample_conv(x, w, b)):
        y1, y2 = fn(d1), fn(d2)
        assert hasattr(y1, "_glare_gn_stats") and not hasattr(y2, "_glare_gn_stats")
        assert torch.equal(y1, y2)
        ref_stats = d2.ops.gn_stats(y2.permute(0, 2, 3, 1).contiguous(), B, y2.shape[2] * y2.shape[3], C)
        assert torch.allclose(y1._glare_gn_stats, ref_stats, rtol=1e-5, atol=1e-3)
        assert torch.equal(fn(d1)._glare_gn_stats, y1._glare_gn_stats)          # deterministic: fixed reduction order, no atomics
        o1, o2 = d1.gn_swish(y1, gamma, beta), d2.gn_swish(y2, gamma, beta)
        # the two statistics agree to ~1e-7 relative: at most one unit in the last place of the operand (8 bits for mode 0, 16 for mode 4)
        assert float((o1.dense() - o2.dense()).abs().max()) < {0: 1e-2, 4: 6.2e-5}.get(mode, 2e-5)


def test_cat_operand_equals_cat_then_convert(glare_lib):
    """WarpBlock's cat([x_vq, h]) written directly as the offset conv's operand: bitwise the operand of the concatenated fp32 tensor"""
    from glare_b200 import ops
    from glare_b200.dense import TcDense
    g = torch.Generator().manual_seed(12)
    a = torch.randn((2, 128, 13, 21), generator=g).cuda().contiguous(memory_format=torch.channels_last)
    b = torch.randn((2, 256, 13, 21), generator=g).cuda().contiguous(memory_format=torch.channels_last)
    d = TcDense(4)
    op = d.cat_operand(a, b)
    want, _ = ops.conv_prep_act(4, torch.cat([a, b], dim=1).permute(0, 2, 3, 1).contiguous())
    assert op.C == 384 and torch.equal(op.hi.view(torch.int16), want.view(torch.int16))
    assert TcDense(0).cat_operand(a, b) is None


def test_attention_pv_key_bands(glare_lib):
    """P V GEMM contracted in key bands chained through the residual epilogue (1080p path) == the single launch up to fp32 rounding of the band
    sums, for the fp32 output and for the operand-packing output; ragged last band"""
    from glare_b200.dense import TcDense, Operand
    B, C, h, w = 2, 512, 24, 21                      # 504 keys -> Np = 512: bands of 192 -> 192 + 192 + 128
    g = torch.Generator().manual_seed(21)
    q, k, v = (torch.randn((B, C, h, w), generator=g).cuda() for _ in range(3))
    one, banded = TcDense(4), TcDense(4)
    banded.attn_key_band = 192
    o1, o2 = one.attention(q, k, v), banded.attention(q, k, v)
    ref = _attention_fp64(q, k, v)
    sc = max(1.0, float(ref.abs().max()))
    assert float((o2.reshape(B, C, h * w).double() - ref).abs().max()) < 6e-5 * sc
    assert float((o1 - o2).abs().max()) < 2e-5 * sc
    p1, p2 = one.attention(q, k, v, as_operand=True), banded.attention(q, k, v, as_operand=True)
    assert isinstance(p2, Operand) and float((p1.dense() - p2.dense()).abs().max()) < 3e-5 * sc
    b0 = TcDense(0)
    b0.attn_key_band = 192
    assert float((b0.attention(q, k, v).reshape(B, C, h * w).double() - ref).abs().max()) < 3e-2 * sc
