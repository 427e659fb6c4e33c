"""CPU: the oracle restatement (oracle/) against the golden vectors generated from the reference's own
modules (oracle/gen_golden.py, tests/golden/PIN_REPORT.txt).  Also re-pins the goldens to the synthetic
checkpoint generator through the stored fingerprints."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from glare_b200 import synth
from oracle import glare_oracle as O
from oracle import vq_lookup


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_fingerprints_match_goldens(sd_g, sd_v):
    g = load_golden("pipe_32x48")
    assert synth.state_fingerprint(sd_g) == pytest.approx(float(g["fingerprint_netG"]), rel=1e-12)
    assert synth.state_fingerprint(sd_v) == pytest.approx(float(g["fingerprint_vqgan"]), rel=1e-12)


def test_vq_indices_bit_exact(sd_v):
    g = load_golden("vq")
    cb = sd_v["quantize.embedding.weight"].numpy()
    idx, zq = vq_lookup(g["z"], cb)
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(zq.view(np.uint32), g["zq"].view(np.uint32))      # straight-through z + (e - z) bits
    idx_l, _ = vq_lookup(g["z_large"], cb)
    assert np.array_equal(idx_l, g["idx_large"].astype(np.int64))


def test_vq_first_index_tie_break(sd_v):
    g = load_golden("vq")
    cb = sd_v["quantize.embedding.weight"].clone()
    cb[4096:] = cb[:4096]
    idx, _ = vq_lookup(g["z"], cb.numpy())
    assert np.array_equal(idx, g["idx_dup"]) and idx.max() < 4096


@pytest.mark.parametrize("s", [0, 2, 15, 27])
def test_flow_step_both_directions(sd_g, s):
    g = load_golden("flow")
    z, ft = T(g["z"]), T(g["ft"])
    coupling = s not in O.NO_COUPLING_STEPS
    p = "flowUpsamplerNet.layers.%d" % s
    zi = O.flow_step_inverse(sd_g, p, z, ft, coupling)
    zf, ld = O.flow_step_forward(sd_g, p, z, ft, torch.zeros(2), coupling)
    assert torch.allclose(zi, T(g["inv_%d" % s]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(zf, T(g["fwd_%d" % s]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(ld, T(g["logdet_%d" % s]), atol=1e-3, rtol=1e-5)


def test_flow_chains_and_round_trip(sd_g):
    g = load_golden("flow")
    z, ft = T(g["z"]), T(g["ft"])
    x = O.flow_decode(sd_g, z, ft)
    assert torch.allclose(x, T(g["decode"]), atol=1e-4, rtol=1e-5)
    zz, ld = O.flow_encode(sd_g, z, ft)
    assert torch.allclose(zz, T(g["encode"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(ld, T(g["encode_logdet"]), atol=1e-2, rtol=1e-5)
    back, _ = O.flow_encode(sd_g, x, ft)                       # invertibility (SURVEY.md 3.4 known answer)
    assert float((back - z).abs().max()) < 5e-4


def test_dcn_matches_reference_kernel_semantics():
    g = load_golden("dcn")
    y = O.modulated_deform_conv(T(g["x"]), T(g["offset"]), T(g["mask"]), T(g["weight"]), T(g["bias"]))
    assert torch.allclose(y, T(g["y"]), atol=1e-5, rtol=1e-5)


def test_taming_blocks(sd_g, sd_v):
    g = load_golden("blocks")
    assert torch.allclose(O.resnet_block(sd_g, "RRDB.encoder.down.0.block.0", T(g["x128"])), T(g["res128"]), atol=1e-5)
    assert torch.allclose(O.resnet_block(sd_g, "RRDB.encoder.down.1.block.0", T(g["x128b"])), T(g["res128_256"]), atol=1e-5)
    assert torch.allclose(O.attn_block(sd_g, "RRDB.encoder.mid.attn_1", T(g["x512"])), T(g["attn512"]), atol=1e-5)
    assert torch.allclose(O.downsample(sd_g, "RRDB.encoder.down.0.downsample", T(g["x128"])), T(g["down128"]), atol=1e-5)
    assert torch.allclose(O.upsample(sd_v, "decoder.up.2.upsample", T(g["x512"]))[:, ::8], T(g["up512"]), atol=1e-5)


def test_pipeline_32x48(sd_g, sd_v):
    g = load_golden("pipe_32x48")
    st = {}
    out = O.glare_infer(sd_g, sd_v, T(g["lr"]), stages=st)
    assert torch.allclose(st["cond_feat"], T(g["cond_feat"]), atol=1e-5)
    assert torch.allclose(st["z_flow"], T(g["z_flow"]), atol=1e-4, rtol=1e-5)
    assert np.array_equal(st["idx"].numpy(), g["idx"].astype(np.int64).reshape(-1))
    assert torch.allclose(out, T(g["out"]), atol=1e-4)
    gt = T(g["gt"])
    assert abs(O.psnr(out.clamp(0, 1), gt) - O.psnr(T(g["out"]).clamp(0, 1), gt)) < 0.01


def test_stage2_gradients_match_reference_autograd():
    """BASELINE config 4: the oracle's stage-2 objective and its autograd gradients against the reference's own
    forward + backward (tests/golden/stage2.npz, written by oracle/gen_golden.py from LLFlowVQGAN2_arch.py:75-122)"""
    from glare_b200 import synth
    g = load_golden("stage2")
    sd2 = {k: v.clone().requires_grad_(True) for k, v in synth.synth_state_dict("netG_stage2", 0).items()}
    z, nll = O.stage2_nll(sd2, T(g["gt_latent"]), T(g["lr"]))
    assert torch.allclose(z, T(g["z"]), atol=1e-4, rtol=1e-5)
    assert torch.allclose(nll, T(g["nll"]), atol=1e-4, rtol=1e-5)
    nll.mean().backward()
    for key in list(g):
        if key.startswith("grad."):
            ref = T(g[key])
            got = sd2[key[5:]].grad
            assert float((got - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), key


def test_flow_explicit_backward_matches_autograd(sd_g):
    """oracle/flow_backward.py (the hand-derived backward the training kernels follow) against autograd of the oracle's flow"""
    from oracle import flow_backward as FB
    gen = torch.Generator().manual_seed(4)
    B, h, w = 2, 6, 7
    gt = torch.randn((B, 3, h, w), generator=gen)
    ft = torch.sigmoid(torch.randn((B, 64, h, w), generator=gen))
    mean = torch.randn((B, 3, h, w), generator=gen) * 0.1
    keys = [k for k in sd_g if k.startswith("flowUpsamplerNet.layers.")]
    sd = dict(sd_g)
    leaf = {k: sd_g[k].clone().requires_grad_(True) for k in keys}
    sd.update(leaf)
    gt_a, ft_a, mean_a = (t.clone().requires_grad_(True) for t in (gt, ft, mean))
    z, logdet = O.flow_encode(sd, gt_a, ft_a)
    import math
    pixels = h * w
    logp = (-0.5 * ((z - mean_a) ** 2 + math.log(2 * math.pi))).sum(dim=(1, 2, 3))
    nll = -(logdet + logp) / (math.log(2.0) * pixels)
    nll.mean().backward()
    with torch.no_grad():
        nll_e, z_e, g_gt, g_ft, g_mean, grads = FB.nll_forward_backward(sd_g, gt, ft, mean)
    assert torch.allclose(nll_e, nll.detach(), atol=1e-4, rtol=1e-5) and torch.allclose(z_e, z.detach(), atol=1e-4, rtol=1e-5)

    def close(a, b, what):
        sc = max(float(b.abs().max()), 1e-6)
        assert float((a - b).abs().max()) <= 2e-4 * sc + 1e-7, (what, float((a - b).abs().max()), sc)

    close(g_gt, gt_a.grad, "gt")
    close(g_ft, ft_a.grad, "ft")
    close(g_mean, mean_a.grad, "mean")
    checked = 0
    for k in keys:
        if leaf[k].grad is None:
            continue                                  # the unused `f` nets of the noCoupling steps
        assert k in grads, k
        close(grads[k], leaf[k].grad, k)
        checked += 1
    assert checked >= 24 * 2 * 9 + 28 * 3


def test_oracle_reproduces_reference_at_bench_shape(sd_g, sd_v):
    """420x620 (BASELINE configs[1] shape, image 0 of the bench batch): oracle against the stored reference outputs"""
    g = load_golden("pipe_420x620")
    lq, _ = synth.synth_images(1, 400, 600, seed=0)
    lr = synth.preprocess(synth.pad_lol(lq))
    assert float(lr.double().sum()) == pytest.approx(float(g["lr_checksum"]), rel=1e-12)
    st = {}
    out = O.glare_infer(sd_g, sd_v, lr, stages=st)
    assert np.array_equal(st["idx"].reshape(-1).numpy(), g["idx"].astype(np.int64))
    assert float((st["z_flow"] - T(g["z_flow"])).abs().max()) <= 1e-5
    assert float((out - T(g["out"])).abs().max()) <= 1e-5
