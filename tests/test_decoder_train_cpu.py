"""CPU: stage-3 training of the deformable decoder (glare_b200/decoder_train.py: tape over MultiScaleDecoder2, DCN backward inside the loop)
with torch restatements of the kernel-level primitives, against torch autograd of the oracle for EVERY parameter the forward uses."""
import numpy as np
import pytest
import torch

from encoder_train_emu import TorchLeaves


def _tv_dcn(x, offset, mask, weight, bias, stride=1, padding=1, dilation=1, dg=4):
    from torchvision.ops import deform_conv2d
    return deform_conv2d(x, offset, weight, bias, stride=stride, padding=padding, dilation=dilation, mask=mask)


def _inputs(B=2, h=4, w=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(s, generator=g)                         # noqa: E731
    return r(B, 3, h, w), [r(B, 256, 2 * h, 2 * w), r(B, 128, 4 * h, 4 * w)], {1: r(B, 256, 2 * h, 2 * w), 0: r(B, 128, 4 * h, 4 * w)}


@pytest.mark.parametrize("global_ratio", [True, False])
def test_decoder_backward_matches_autograd_of_the_oracle(sd_g, global_ratio):
    from glare_b200 import decoder_train
    from oracle import glare_oracle as O
    sd = {k: v.clone() for k, v in sd_g.items() if k.startswith("deformable_decoder.")}
    g = torch.Generator().manual_seed(5)
    for k in sd:                                                         # conv_offset is zero-initialised (deform_conv.py:367-371): give the
        if "conv_offset" in k:                                           # offsets / masks a real dependence on the features
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
    z, vq, mid = _inputs()
    with torch.no_grad():
        tr = decoder_train.DecoderTrainer(TorchLeaves(), sd, global_ratio=global_ratio)
        rec = tr.forward(z, vq, mid)
        g_rec = torch.randn(rec.shape, generator=g)
        grads = tr.backward(g_rec)
    sda = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rec_a = O.aft_decoder(sda, z, vq, mid, per_sample_ratio=not global_ratio, dcn=_tv_dcn)
    assert float((rec - rec_a.detach()).abs().max()) <= 1e-4 * max(1.0, float(rec_a.detach().abs().max()))
    rec_a.backward(g_rec)
    checked = 0
    for k, v in sda.items():
        if v.grad is None:                                               # conv_out, scale.*, bias.*, enc.*: built but unused by the forward
            assert k not in grads, k
            continue
        assert k in grads and grads[k].shape == v.shape, k
        sc = max(float(v.grad.abs().max()), 1e-6)
        assert float((grads[k] - v.grad).abs().max()) <= 1e-3 * sc + 2e-6, (k, float((grads[k] - v.grad).abs().max()), sc)   # k.bias: exactly 0 in theory
        checked += 1
    assert checked > 150 and "deformable_decoder.mix.0.w" in grads and "deformable_decoder.warp.1.dcn.weight" in grads


def test_decoder_autograd_node_fills_param_grads(sd_g):
    """DeformableDecoderFn: the reconstruction carries the graph edge to the parameters; a torch loss on top of it backpropagates like the
    reference's total_loss.backward() (VQLLFLOWD_model.py:226)"""
    from glare_b200 import decoder_train
    from oracle import glare_oracle as O
    params = {k: torch.nn.Parameter(v.clone()) for k, v in sd_g.items() if k.startswith("deformable_decoder.")}
    z, vq, mid = _inputs(B=1, seed=3)
    rec = decoder_train.deformable_decoder(params.items(), z, vq, mid, TorchLeaves())
    gt = torch.rand(rec.shape, generator=torch.Generator().manual_seed(1))
    (rec.clamp(0, 1) - gt).abs().mean().backward()
    sda = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    (O.aft_decoder(sda, z, vq, mid, per_sample_ratio=False, dcn=_tv_dcn).clamp(0, 1) - gt).abs().mean().backward()
    for k in ("deformable_decoder.conv_in.weight", "deformable_decoder.up.0.block.2.conv2.weight", "deformable_decoder.warp.0.offset.weight",
              "deformable_decoder.residual_conv.bias", "deformable_decoder.mix.1.w"):
        ref = sda[k].grad
        assert float((params[k].grad - ref).abs().max()) <= 1e-3 * float(ref.abs().max()) + 1e-8, k
    assert params["deformable_decoder.conv_out.weight"].grad is None
    # loss scaling (GradScaler, VQLLFLOWD_model.py:226) is linear in the gradients; a second backward through the freed tape is refused
    for p in params.values():
        p.grad = None
    rec = decoder_train.deformable_decoder(params.items(), z, vq, mid, TorchLeaves())
    loss = (rec.clamp(0, 1) - gt).abs().mean() * 1024.0
    loss.backward(retain_graph=True)
    k = "deformable_decoder.up.0.block.2.conv2.weight"
    assert float((params[k].grad / 1024.0 - sda[k].grad).abs().max()) <= 1e-3 * float(sda[k].grad.abs().max())
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()


def test_stage3_evaluation_matches_the_reference_golden(sd_g, sd_v):
    """one stage-3 objective + gradient evaluation (VQLLFLOWD_model.py:187-232) against the unmodified reference's own numbers
    (tests/golden/stage3.npz, oracle/gen_golden_stage3.py): the decoder tape on torch leaves, csrc/loss.cu on the host, frozen stages from
    the oracle"""
    from conftest import load_golden
    from glare_b200 import decoder_train, losses, synth
    from oracle import glare_oracle as O
    from oracle.gen_golden_stage3 import vgg_state
    from test_losses_cpu import _host_kernels
    g = load_golden("stage3")
    lq, gt = torch.from_numpy(g["lq"]), torch.from_numpy(g["gt"])
    with torch.no_grad():
        st = {}
        O.glare_infer(sd_g, sd_v, synth.preprocess(lq), per_sample_ratio=False, stages=st)
    assert float((st["z_flow"] - torch.from_numpy(g["z_flow"])).abs().max()) < 1e-4
    params = {k: torch.nn.Parameter(v.clone()) for k, v in sd_g.items() if k.startswith("deformable_decoder.")}
    rec = decoder_train.deformable_decoder(params.items(), st["z_flow"], [st["vq_feat1"], st["vq_feat0"]], {1: st["mid1"], 0: st["mid0"]},
                                           TorchLeaves(), global_ratio=True)
    assert float((rec.detach() - torch.from_numpy(g["rec"])).abs().max()) < 2e-4
    K = _host_kernels()
    percep = losses.PerceptualNetwork(state_dict=vgg_state(0), leaves=TorchLeaves())
    total, terms = losses.stage3_loss(rec, gt, percep, msssim_fn=lambda x, y, **kw: losses.msssim(x, y, kernels=K, **kw))
    for k, v in terms.items():
        assert abs(float(v.detach()) - float(g["term." + k])) <= 2e-5 * max(abs(float(g["term." + k])), 1e-3), k
    total.backward()
    with_grad = set(g["with_grad"].tolist())
    assert {k for k, p in params.items() if p.grad is not None} == with_grad
    gmax = max(float(np.abs(g[k]).max()) for k in g if k.startswith("grad."))
    for k in g:
        if k.startswith("grad."):
            ref = torch.from_numpy(g[k])
            err = float((params[k[5:]].grad - ref).abs().max())
            assert err <= 2e-3 * max(float(ref.abs().max()), 1e-4 * gmax), (k, err, float(ref.abs().max()))
    for k, s in zip(g["with_grad"].tolist(), g["abs_sum"].tolist()):         # every parameter: checksum of the gradient
        got = float(params[k].grad.double().abs().sum())
        assert abs(got - s) <= 5e-3 * max(s, 1e-4 * gmax * params[k].numel() ** 0.5), (k, got, s)
