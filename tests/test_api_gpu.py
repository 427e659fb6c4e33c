"""GPU: the public entry (GlareEnhancer.enhance: host uint8 in -> host uint8 out) against the oracle run through the
reference's own pre/post-processing steps (infer_dataset_lol.py:124-135 'lol' padding, infer_unpaired.py:81-88,130 'auto')."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_u8(sd_g, sd_v, lq_u8, pad):
    from glare_b200 import synth
    from oracle import glare_oracle as O
    x = lq_u8.permute(0, 3, 1, 2).float() / 255.0
    h, w = x.shape[-2:]
    if pad == "lol":
        xp = synth.pad_lol(x)
        box = (0, h, 20, 20 + w)
    else:
        xp, (h1, h2, w1, w2) = synth.auto_padding(x)
        box = (h1, h1 + h, w1, w1 + w)
    out = O.glare_infer(sd_g, sd_v, O.preprocess(xp))
    out = out[:, :, box[0]:box[1], box[2]:box[3]].clamp(0, 1) * 255.0
    return out.to(torch.uint8).permute(0, 2, 3, 1)


@pytest.mark.parametrize("pad,hw", [("lol", (44, 60)), ("auto", (40, 56)), ("auto", (32, 48))])
def test_enhance_matches_oracle(glare_lib, sd_g, sd_v, pad, hw):
    from glare_b200 import synth
    from glare_b200.api import GlareEnhancer
    lq, _ = synth.synth_images(2, hw[0], hw[1], seed=9)
    lq_u8 = (lq.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous()
    enh = GlareEnhancer(sd_g, sd_v, device="cuda:0", pad=pad)
    out = enh.enhance(lq_u8.pin_memory())
    ref = _oracle_u8(sd_g, sd_v, lq_u8, pad)
    assert out.shape == ref.shape == lq_u8.shape and out.dtype == torch.uint8
    diff = (out.int() - ref.int()).abs()
    # uint8 truncation: a 1e-3 float difference can move a value across an integer boundary, never by more than 1
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 0.02


def test_enhance_rejects_wrong_input(glare_lib, sd_g, sd_v):
    from glare_b200.api import GlareEnhancer
    enh = GlareEnhancer(sd_g, sd_v, device="cuda:0")
    with pytest.raises(ValueError):
        enh.enhance(torch.zeros((1, 3, 8, 8)))


@pytest.mark.parametrize("pad,hw", [("lol", (37, 53)), ("auto", (33, 47)), ("auto", (32, 48)), ("lol", (400, 600))])
def test_pre_and_postprocess_kernels(glare_lib, pad, hw):
    """glare_preprocess_u8 / glare_postprocess_u8 against the reference steps restated in glare_b200.synth (np.pad 'reflect',
    cv2.BORDER_REFLECT, /255, log(clamp), clip*255 -> uint8 truncation)"""
    from glare_b200 import ops, synth
    g = torch.Generator().manual_seed(hw[0])
    u8 = torch.randint(0, 256, (2, hw[0], hw[1], 3), generator=g, dtype=torch.uint8)
    x = u8.permute(0, 3, 1, 2).float() / 255.0
    h, w = hw
    if pad == "lol":
        ref, p, mode, box = synth.preprocess(synth.pad_lol(x)), (0, 20, 20, 0), 0, (0, h, 20, 20 + w)
    else:
        xp, (h1, h2, w1, w2) = synth.auto_padding(x)
        ref, p, mode, box = synth.preprocess(xp), (h1, h2, w1, w2), 1, (h1, h1 + h, w1, w1 + w)
    lr = ops.preprocess_u8(u8.cuda(), p, mode)
    assert lr.shape == ref.shape and float((lr.cpu() - ref).abs().max()) < 2e-6
    y = torch.randn(ref.shape, generator=g) * 0.6 + 0.5
    want = (y[:, :, box[0]:box[1], box[2]:box[3]].clamp(0, 1) * 255.0).to(torch.uint8).permute(0, 2, 3, 1)
    got = ops.postprocess_u8(y.cuda().contiguous(memory_format=torch.channels_last), box).cpu()
    assert torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
def test_aft_axpby_matches_reference_rounding(glare_lib, layout):
    """Mix.forward (deformableDecoder_arch.py:587-590) and the mean-ratio residual (:567) in one pass each: bitwise what the
    reference's separate mul / mul / add kernels produce"""
    from glare_b200 import ops
    g = torch.Generator().manual_seed(2)
    a = torch.randn((3, 128, 17, 22), generator=g).cuda()
    b = torch.randn((3, 128, 17, 22), generator=g).cuda()
    if layout == "nhwc":
        a, b = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
    m = torch.sigmoid(torch.tensor([0.3])).cuda()
    assert torch.equal(ops.aft_axpby(a, b, m, 1 - m), a * m + b * (1 - m))
    ratio = (a.mean(dim=(1, 2, 3), keepdim=True) / b.mean(dim=(1, 2, 3), keepdim=True))
    one = torch.ones((1,), device="cuda")
    assert torch.equal(ops.aft_axpby(a, b, one, ratio), a + b * ratio)
    with pytest.raises(ValueError):
        ops.aft_axpby(a, b[:, :64], m, m)


def test_graph_replay_equals_eager(glare_lib, sd_g, sd_v):
    """the whole pass as ONE captured CUDA graph (engine.graphed): replays on changing inputs give what the eager launches give, one
    graph per input shape, and the launch counter advances by the graph's kernel count per replay"""
    from glare_b200 import ops, synth
    from glare_b200.api import GlareEnhancer
    enh = GlareEnhancer(sd_g, sd_v, device="cuda:0", pad="lol")
    batches = []
    for seed in (1, 2):
        lq, _ = synth.synth_images(2, 44, 60, seed=seed)
        batches.append((lq.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory())
    for i, x in enumerate((batches[0], batches[1], batches[0])):
        n0 = ops.LAUNCHES
        got = enh.enhance(x, graph=True).clone()
        n_graph = ops.LAUNCHES - n0
        n0 = ops.LAUNCHES
        want = enh.enhance(x, graph=False)
        n_eager = ops.LAUNCHES - n0
        diff = (got.int() - want.int()).abs()
        assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 1e-3
        if i > 0:                                # (the first call also counts the eager warm-up and the capture pass)
            assert n_graph == n_eager > 300      # kernels inside the graph are counted per replay
    assert len(enh.engine._graphs) == 1
    st = {}
    lr = synth.preprocess(synth.synth_images(1, 32, 48, seed=3)[0])
    a = enh.engine.infer(lr, stages=st, graph=True).clone()
    b = enh.engine.infer(lr, graph=False)
    assert torch.allclose(a, b, atol=1e-5) and st["idx"].numel() == 8 * 12


@pytest.mark.parametrize("graph", [True, False])
def test_attention_flag_trip_recaptures_and_recomputes(glare_lib, sd_g, sd_v, graph):
    """a fused-softmax row outside its window (forced here by a row reference 300 below the sampled maximum: every p~ overflows) raises the
    device flag; the engine -- graph replay or eager -- switches the backend to the exact softmax, drops / re-captures the graph and
    returns the recomputed result, equal to an engine that never used the fused path"""
    from glare_b200 import synth
    from glare_b200.dense import make_dense
    from glare_b200.engine import GlareEngine
    lr = synth.preprocess(synth.synth_images(2, 32, 48, seed=6)[0])
    bad = make_dense("auto")
    bad.attn_ref_offset = -300.0
    eng = GlareEngine(sd_g, sd_v, device="cuda:0", dense=bad)
    out = eng.infer(lr, graph=graph).clone()
    assert not bad.attn_fused and bad.fallbacks
    exact = make_dense("auto")
    exact.attn_fused = False
    ref = GlareEngine(sd_g, sd_v, device="cuda:0", dense=exact).infer(lr)
    assert torch.isfinite(out).all() and float((out - ref).abs().max()) < 1e-5
    again = eng.infer(lr, graph=graph)              # stays on the exact path, one graph for the shape
    assert float((again - ref).abs().max()) < 1e-5
    if graph:
        assert len(eng._graphs) == 1
