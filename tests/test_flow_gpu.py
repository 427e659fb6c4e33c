"""GPU parity: fused flow-step kernels through the C ABI vs golden vectors (reference FlowStep outputs) and the
oracle.  Tolerance (fp32 path): 2e-4 abs + 2e-5 rel on z (|z| <= ~80 after 28 inverse steps); logdet 1e-4 rel."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu
ATOL, RTOL = 2e-4, 2e-5


def _conv(x, w):
    torch.backends.cudnn.allow_tf32 = False
    return F.conv2d(x, w, None, padding=1)


@pytest.fixture(scope="module")
def plan(glare_lib, sd_g):
    from glare_b200 import flow
    return flow.FlowPlan(sd_g, torch.device("cuda:0"))


def _single_step(plan, s, z, ft, direction):
    """one step through glare_flow_step_f32 (plus the batched NN_F tail), as FlowStep.forward would"""
    from glare_b200 import flow, ops
    B, _, h, w = z.shape
    hw, n = h * w, len(flow.COUPLING_STEPS)
    out = torch.empty_like(z)
    ld = torch.zeros(B, device=z.device)
    pw = (plan.pw_inv if direction == 1 else plan.pw_fwd)[s]
    if s in flow.NO_COUPLING_STEPS:
        ops.flow_step(direction, False, z, out, None, (0, 0, 0), None, 0, None, pw, None)
    else:
        P, hF = flow.precompute(plan, ft, _conv)
        ci = flow.COUPLING_STEPS.index(s)
        ops.flow_step(direction, True, z, out, P[:, ci * 128:], flow.plane_strides(P, False), hF[:, ci * 6:], n * 6 * hw, plan.nets_a[ci], pw, ld)
    sign = -1.0 if direction == 1 else 1.0
    ld = ld + sign * (plan.ld_const[s, 0] + plan.ld_const[s, 1]) * float(hw)
    torch.cuda.synchronize()
    return out, ld


@pytest.mark.parametrize("s", [0, 2, 15, 27])
def test_step_against_golden(plan, s):
    g = load_golden("flow")
    z, ft = torch.from_numpy(g["z"]).cuda(), torch.from_numpy(g["ft"]).cuda()
    zi, _ = _single_step(plan, s, z, ft, 1)
    zf, ld = _single_step(plan, s, z, ft, 0)
    assert torch.allclose(zi.cpu(), torch.from_numpy(g["inv_%d" % s]), atol=ATOL, rtol=RTOL)
    assert torch.allclose(zf.cpu(), torch.from_numpy(g["fwd_%d" % s]), atol=ATOL, rtol=RTOL)
    assert torch.allclose(ld.cpu(), torch.from_numpy(g["logdet_%d" % s]), atol=1e-2, rtol=1e-4)


def test_chains_against_golden(plan):
    from glare_b200 import flow
    g = load_golden("flow")
    z, ft = torch.from_numpy(g["z"]).cuda(), torch.from_numpy(g["ft"]).cuda()
    x, _ = flow.decode(plan, z, ft, _conv)
    ref = torch.from_numpy(g["decode"])
    assert torch.allclose(x.cpu(), ref, atol=2e-3, rtol=1e-4), float((x.cpu() - ref).abs().max())
    zz, ld = flow.encode(plan, z, ft, _conv)
    assert torch.allclose(zz.cpu(), torch.from_numpy(g["encode"]), atol=ATOL, rtol=RTOL)
    assert torch.allclose(ld.cpu(), torch.from_numpy(g["encode_logdet"]), atol=5e-2, rtol=1e-4)


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 13, 29), (3, 14, 14), (2, 15, 28), (1, 105, 155)])
def test_every_step_against_oracle_ragged_tiles(plan, sd_g, shape):
    """all 28 steps, both directions, on shapes that straddle the 14x14 tiling (edges, single pixel)"""
    from oracle import glare_oracle as O
    B, h, w = shape
    g = torch.Generator().manual_seed(h * 100 + w)
    z = torch.randn((B, 3, h, w), generator=g)
    ft = torch.sigmoid(torch.randn((B, 64, h, w), generator=g))
    steps = range(28) if h < 100 else (2, 27)
    for s in steps:
        coupling = s not in O.NO_COUPLING_STEPS
        p = "flowUpsamplerNet.layers.%d" % s
        zi_o = O.flow_step_inverse(sd_g, p, z, ft, coupling)
        zf_o, ld_o = O.flow_step_forward(sd_g, p, z, ft, torch.zeros(B), coupling)
        zi, _ = _single_step(plan, s, z.cuda(), ft.cuda(), 1)
        zf, ld = _single_step(plan, s, z.cuda(), ft.cuda(), 0)
        assert torch.allclose(zi.cpu(), zi_o, atol=ATOL, rtol=RTOL), (s, float((zi.cpu() - zi_o).abs().max()))
        assert torch.allclose(zf.cpu(), zf_o, atol=ATOL, rtol=RTOL), (s, float((zf.cpu() - zf_o).abs().max()))
        assert torch.allclose(ld.cpu(), ld_o, atol=1e-3 * h * w + 1e-3, rtol=1e-4), s


def test_full_size_round_trip_and_logdet(plan):
    """BASELINE config-2 size: decode(encode(x)) == x and logdet(encode) == -logdet(decode) (size-independent)"""
    from glare_b200 import flow
    g = torch.Generator().manual_seed(3)
    B, h, w = 4, 105, 155
    x = torch.randn((B, 3, h, w), generator=g).cuda()
    ft = torch.sigmoid(torch.randn((B, 64, h, w), generator=g)).cuda()
    z, ld = flow.encode(plan, x, ft, _conv)
    ld2 = torch.zeros(B, device="cuda")
    back, ld2 = flow.decode(plan, z, ft, _conv, logdet=ld2)
    assert float((back - x).abs().max()) < 2e-3
    assert torch.allclose(ld, -ld2, rtol=1e-4, atol=1.0)


def test_empty_batch(plan):
    from glare_b200 import ops
    z = torch.zeros((0, 3, 8, 8), device="cuda")
    out = torch.empty_like(z)
    ops.flow_step(1, False, z, out, None, (0, 0, 0), None, 0, None, plan.pw_inv[0], None)


def test_aliasing_is_rejected(plan):
    from glare_b200 import ops
    z = torch.zeros((1, 3, 8, 8), device="cuda")
    with pytest.raises(RuntimeError):
        ops.flow_step(1, False, z, z, None, (0, 0, 0), None, 0, None, plan.pw_inv[0], None)
