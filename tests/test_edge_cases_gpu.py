"""GPU: degenerate inputs at the operator boundary -- empty batches (every entry point returns without launching and the wrappers hand back
empty tensors of the right shape), single-pixel / single-token shapes, and maximum-magnitude latents for the VQ lookup."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_empty_batches(glare_lib, sd_g, sd_v):
    from glare_b200 import ops
    from glare_b200.dense import make_dense
    dev = torch.device("cuda:0")
    cb = ops.vq_pack_codebook(sd_v["quantize.embedding.weight"].to(dev))
    idx, zq = ops.vq_lookup(torch.empty((0, 3, 5, 7), device=dev), cb)
    assert idx.shape == (0,) and zq.shape == (0, 3, 5, 7)
    w = torch.randn((128, 128, 3, 3), device=dev) * 0.03
    y = ops.modulated_deform_conv(torch.empty((0, 128, 6, 6), device=dev), torch.empty((0, 72, 6, 6), device=dev),
                                  torch.empty((0, 36, 6, 6), device=dev), w, None, 1, 1, 1, 1, 4)
    assert y.shape == (0, 128, 6, 6)
    d = make_dense("auto")
    y = d.conv2d(torch.empty((0, 128, 8, 16), device=dev), w)
    assert y.shape == (0, 128, 8, 16)
    op = d.gn_swish(torch.empty((0, 128, 8, 16), device=dev), torch.ones(128, device=dev), torch.zeros(128, device=dev))
    assert op.B == 0
    torch.cuda.synchronize()


def test_single_token_and_tiny_images(glare_lib, sd_g, sd_v):
    """the smallest inputs the path accepts: one latent token for the VQ, a 16 x 16 image (4 x 4 latent) end to end against the oracle"""
    from glare_b200 import ops, synth
    from glare_b200.dense import make_dense
    from glare_b200.engine import GlareEngine
    from oracle import glare_oracle as O
    from oracle import vq_lookup
    dev = torch.device("cuda:0")
    cbw = sd_v["quantize.embedding.weight"]
    z = torch.tensor([[[[0.3]], [[-1.2]], [[0.7]]]])
    idx, zq = ops.vq_lookup(z.to(dev), ops.vq_pack_codebook(cbw.to(dev)))
    idx_o, zq_o = vq_lookup(z.numpy(), cbw.numpy())
    assert int(idx.cpu()) == int(idx_o[0]) and np.array_equal(zq.cpu().numpy().view(np.uint32), zq_o.view(np.uint32))
    lq, gt = synth.synth_images(1, 16, 16, seed=2)
    lr = synth.preprocess(lq)
    eng = GlareEngine(sd_g, sd_v, device="cuda:0", dense=make_dense("auto"))
    st = {}
    out = eng.infer(lr, stages=st).cpu()
    st_o = {}
    ref = O.glare_infer(sd_g, sd_v, lr, stages=st_o)
    assert float((st["z_flow"].cpu() - st_o["z_flow"]).abs().max()) < 5e-3
    if bool((st["idx"].cpu() == st_o["idx"]).all()):
        assert float((out - ref).abs().max()) < 1e-3


def test_vq_extreme_latents(glare_lib, sd_v):
    """largest finite magnitudes, infinities and NaN tokens: indices equal the C oracle's (which restates the reference's fp32 recipe)"""
    from glare_b200 import ops
    from oracle import vq_lookup
    cbw = sd_v["quantize.embedding.weight"]
    vals = [3.0e38, -3.0e38, 1e19, -1e19, 1e-38, 0.0, -0.0, float("inf"), float("-inf"), float("nan")]
    g = torch.Generator().manual_seed(0)
    z = torch.tensor(vals)[torch.randint(0, len(vals), (2, 3, 9, 11), generator=g)]
    z[0, :, :4] = torch.randn((3, 4, 11), generator=g)
    idx, _ = ops.vq_lookup(z.cuda(), ops.vq_pack_codebook(cbw.cuda()))
    idx_o, _ = vq_lookup(z.numpy(), cbw.numpy())
    assert np.array_equal(idx.cpu().numpy(), idx_o)


def _ref_ext():
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
    return ref_ext


def test_dcn_border_and_nonfinite_offsets_against_the_reference_kernel(glare_lib):
    """sampling positions exactly on the open border (-1, H) / (-1, W) (deform_conv_cuda_kernel.cu:618), on integer grid points, far outside,
    infinite and NaN: both glare DCN kernels against the reference's own CUDA kernel (built unmodified for sm_100)"""
    from glare_b200 import ops
    ref_ext = _ref_ext()
    g = torch.Generator().manual_seed(17)
    B, C, H, W, dg = 1, 128, 12, 20, 4
    x = torch.randn((B, C, H, W), generator=g).cuda()
    w = (torch.randn((C, C, 3, 3), generator=g) / (3 * C ** 0.5)).cuda()
    b = torch.randn((C,), generator=g).cuda()
    off = torch.randn((B, dg * 18, H, W), generator=g) * 1.5
    special = torch.tensor([0.0, -1.0, 1.0, 0.5, -0.5, -2.0, 2.0, 1e6, -1e6, 3e38, float("inf"), float("-inf"), float("nan"), 1e-30, 0.999999, -0.999999])
    pick = torch.randint(0, len(special), off.shape, generator=g)
    use = torch.rand(off.shape, generator=g) < 0.5
    off = torch.where(use, special[pick], off).cuda()
    msk = torch.sigmoid(torch.randn((B, dg * 9, H, W), generator=g)).cuda()
    out = x.new_empty((B, C, H, W))
    ref_ext.modulated_deform_conv_forward(x, w, b, x.new_empty(0), off, msk, out, x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, dg, True)
    finite = torch.isfinite(out)
    assert float(finite.float().mean()) > 0.5            # (NaN offsets poison the reference's output pixels too: compare where it is finite)
    y_tc = ops.modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, dg)                                   # tensor-core kernel (GLARE's shape)
    y_fma = ops.modulated_deform_conv(x, off, msk, w, b, 1, 1, 1, 1, dg, packed_weight=ops.dcn_pack_weight(w))   # fp32 FMA kernel
    for name, y in (("tc", y_tc), ("fma", y_fma)):
        assert bool((torch.isfinite(y) == finite).all()), name
        d = float((torch.where(finite, y - out, torch.zeros_like(y))).abs().max())
        assert d < 2e-4 * max(1.0, float(out[finite].abs().max())), (name, d)


def test_groupnorm_constant_and_huge_inputs(glare_lib):
    """zero variance (rstd = 1 / sqrt(eps)), a large common offset (fp64 statistics: no E[x^2] - E[x]^2 cancellation) and 1e4-scale values,
    separate statistics kernel and conv-epilogue statistics alike, against torch group_norm in fp64"""
    from glare_b200.dense import TcDense
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(2)
    B, C, H, W = 2, 128, 17, 23
    gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).cuda(), (0.1 * torch.randn(C, generator=g)).cuda()
    cases = {"constant": torch.full((B, C, H, W), 3.25), "offset": 1000.0 + torch.randn((B, C, H, W), generator=g),
             "large": 1e4 * torch.randn((B, C, H, W), generator=g)}
    d = TcDense(2)                                              # 3xTF32 operand: hi + lo is the exact fp32 value
    for name, x in cases.items():
        x = x.cuda()
        got = d.gn_swish(x, gamma, beta, swish=False).dense()
        want = F.group_norm(x.double(), 32, gamma.double(), beta.double(), eps=1e-6).float()
        tol = 2e-3 if name == "offset" else 2e-5                # x - mean itself carries 1e-7 * 1000 of fp32 rounding at a 1000 offset
        assert float((got - want).abs().max()) < tol * max(1.0, float(want.abs().max())), name


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 5), (1, 7, 9), (3, 16, 8), (1, 9, 130)])
def test_dense_ops_on_tiny_and_ragged_spatial_sizes(glare_lib, shape):
    """feature maps smaller than one 8 x 16 pixel tile, odd sizes, a single pixel: 3x3 / 1x1 convs (residual, fused GroupNorm statistics),
    Downsample, Upsample+conv and GroupNorm+swish against cuDNN / torch fp32"""
    import torch.nn.functional as F
    from glare_b200.dense import TcDense
    torch.backends.cudnn.allow_tf32 = False
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 31 + W)
    C = 128
    x = torch.randn((B, C, H, W), generator=g).cuda()
    res = torch.randn((B, C, H, W), generator=g).cuda()
    w3 = (torch.randn((C, C, 3, 3), generator=g) / (3 * C ** 0.5)).cuda()
    w1 = (torch.randn((C, C, 1, 1), generator=g) / C ** 0.5).cuda()
    b = torch.randn((C,), generator=g).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).cuda(), (0.1 * torch.randn(C, generator=g)).cuda()
    d = TcDense(4)

    def close(got, want, what, tol=1e-4):
        assert got.shape == want.shape, what
        assert float((got.float() - want).abs().max()) < tol * max(1.0, float(want.abs().max())), what

    y = d.conv2d(x, w3, b, residual=res)
    close(y, F.conv2d(x, w3, b, padding=1) + res, "conv3x3")
    gn = d.gn_swish(y, gamma, beta)                       # statistics from the conv epilogue
    yn = F.group_norm(y.float(), 32, gamma, beta, eps=1e-6)
    close(gn.dense(), yn * torch.sigmoid(yn), "groupnorm+swish after conv", 2e-4)
    close(d.conv2d(x, w1, b, padding=0), F.conv2d(x, w1, b), "conv1x1")
    if H >= 2 and W >= 2:
        close(d.downsample_conv(x, w3, b), F.conv2d(F.pad(x, (0, 1, 0, 1)), w3, b, stride=2), "downsample")
    close(d.upsample_conv(x, w3, b), F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w3, b, padding=1), "upsample")
