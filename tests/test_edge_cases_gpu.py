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
