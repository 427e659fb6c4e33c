"""Host side of the conditional flow (FlowUpsamplerNet.py:228-326): parameter packing and the 28-step chains.

Parameter names are the reference's state-dict keys (``flowUpsamplerNet.layers.{s}.actnorm.bias`` ...), so a
``net_G.pth`` loads unchanged.  Everything that depends only on the weights -- exp(logs), the fp64 inverse of
the 1x1 conv (Permutations.py:38), slogdet -- is evaluated ONCE here (on the host, in the reference's own
precision) instead of once per call as the reference does, which also removes the per-step host sync
(Permutations.py:25).
"""
import math

import torch

from . import ops

N_FLOW_STEPS = 28
NO_COUPLING_STEPS = (0, 1, 14, 15)          # FlowUpsamplerNet.py:95-106 with LOL.yml K=12, L=2, additionalFlowNoAffine=2
COUPLING_STEPS = tuple(s for s in range(N_FLOW_STEPS) if s not in NO_COUPLING_STEPS)
NET_FLOATS = 9552                           # include/glare_b200.h GLARE_FLOW_NET_FLOATS
HIDDEN = 64


def pack_net(sd, p, has_z):
    """One coupling net (FlowAffineCouplingsAblation.py:143-151) -> packed block, layout in glare_b200.h."""
    w1 = sd[p + ".0.weight"].float().cpu()
    w3 = sd[p + ".4.weight"].float().cpu()
    nout = w3.shape[0]
    out = torch.zeros(NET_FLOATS, dtype=torch.float32)
    if has_z:
        out[0:576] = w1[:, 0].reshape(HIDDEN, 9).flatten()
    out[576:640] = sd[p + ".0.actnorm.bias"].float().cpu().flatten()
    out[640:704] = torch.exp(sd[p + ".0.actnorm.logs"].float().cpu()).flatten()          # FlowActNorms.py:62
    out[704:4800] = sd[p + ".2.weight"].float().cpu()[:, :, 0, 0].t().contiguous().flatten()
    out[4800:4864] = sd[p + ".2.actnorm.bias"].float().cpu().flatten()
    out[4864:4928] = torch.exp(sd[p + ".2.actnorm.logs"].float().cpu()).flatten()
    w3p = torch.zeros(HIDDEN, 9, 8)
    w3p[:, :, :nout] = w3.reshape(nout, HIDDEN, 9).permute(1, 2, 0)
    out[4928:9536] = w3p.flatten()
    out[9536:9536 + nout] = sd[p + ".4.bias"].float().cpu().flatten()
    out[9544:9544 + nout] = torch.exp(sd[p + ".4.logs"].float().cpu() * 3).flatten()       # flow.py:68-70 logscale_factor=3
    return out


def pack_pointwise(sd, p, reverse):
    """ActNorm + InvertibleConv1x1 of one step -> 16 floats (M row-major, bias, scale)."""
    w = sd[p + ".invconv.weight"].float().cpu()
    logs = sd[p + ".actnorm.logs"].float().cpu().flatten()
    out = torch.zeros(16, dtype=torch.float32)
    if reverse:
        out[0:9] = torch.inverse(w.double()).float().flatten()                            # Permutations.py:38
        out[12:15] = torch.exp(-logs)                                                     # FlowActNorms.py:64
    else:
        out[0:9] = w.flatten()
        out[12:15] = torch.exp(logs)
    out[9:12] = sd[p + ".actnorm.bias"].float().cpu().flatten()
    return out


class FlowPlan:
    """Device-resident packed parameters of the whole flow (built once per checkpoint)."""

    def __init__(self, sd, device, prefix="flowUpsamplerNet"):
        self.prefix = prefix
        self.device = device
        nets_a, nets_f, w_a, w_f = [], [], [], []
        for s in COUPLING_STEPS:
            p = "%s.layers.%d.affine" % (prefix, s)
            nets_a.append(pack_net(sd, p + ".fAffine", True))
            nets_f.append(pack_net(sd, p + ".fFeatures", False))
            w_a.append(sd[p + ".fAffine.0.weight"].float().cpu()[:, 1:])      # ft channels of cat([z1, ft]) (:137-141)
            w_f.append(sd[p + ".fFeatures.0.weight"].float().cpu())
        self.nets_a = torch.stack(nets_a).to(device)
        self.nets_f = torch.stack(nets_f).to(device)
        # one dense conv for the ft part of every first layer: rows [ci*128 + 0..63] = NN_A, [ci*128 + 64..127] = NN_F
        self.w_pre = torch.stack([torch.cat([a, f], 0) for a, f in zip(w_a, w_f)]).reshape(-1, HIDDEN, 3, 3).contiguous().to(device)
        self.pw_inv = torch.stack([pack_pointwise(sd, "%s.layers.%d" % (prefix, s), True) for s in range(N_FLOW_STEPS)]).to(device)
        self.pw_fwd = torch.stack([pack_pointwise(sd, "%s.layers.%d" % (prefix, s), False) for s in range(N_FLOW_STEPS)]).to(device)
        # weight-only logdet terms per step (FlowActNorms.py:66-74, Permutations.py:27,51-53), multiplied by `pixels` at run time
        self.ld_const = torch.stack([
            torch.stack([sd["%s.layers.%d.actnorm.logs" % (prefix, s)].float().cpu().sum(),
                         torch.slogdet(sd["%s.layers.%d.invconv.weight" % (prefix, s)].float().cpu())[1]])
            for s in range(N_FLOW_STEPS)]).to(device)


def precompute(plan, ft, conv2d):
    """ft [B,64,h,w] -> (P_all [B, 24*128, h, w], hF_all [B, 24*6, h, w]).
    ``conv2d(x, weight)`` is the dense 3x3 conv path (padding 1, no bias)."""
    B, _, h, w = ft.shape
    n = len(COUPLING_STEPS)
    P = conv2d(ft, plan.w_pre)
    hw = h * w
    hF = torch.empty((B, n * 6, h, w), device=ft.device, dtype=torch.float32)
    ops.flow_cond_tail(P[:, 64:], plane_strides(P, True), plan.nets_f, n, 6, B, h, w, hF, n * 6 * hw, 6 * hw)
    return P, hF


def plane_strides(P, with_step):
    """element strides of the pre-activation tensor P [B, 24*128, h, w] in whatever dense layout the conv path
    produced (NCHW planes or channels_last): (batch, [step,] channel, pixel)"""
    sb, sc, sh, sw = P.stride()
    if sh != sw * P.shape[3]:
        raise ValueError("pre-activation planes must be dense over (h, w)")
    return (sb, 128 * sc, sc, sw) if with_step else (sb, sc, sw)


def decode(plan, z, ft, conv2d, logdet=None, trace=None):
    """FlowUpsamplerNet.decode (FlowUpsamplerNet.py:290-326): steps 27 -> 0.  Returns (x, logdet)."""
    z = z.float().contiguous()
    B, _, h, w = z.shape
    hw, n = h * w, len(COUPLING_STEPS)
    P, hF = precompute(plan, ft, conv2d)
    bufs = [z, torch.empty_like(z), torch.empty_like(z)]
    cur, nxt = 0, 1
    for s in range(N_FLOW_STEPS - 1, -1, -1):
        coupling = s not in NO_COUPLING_STEPS
        if coupling:
            ci = COUPLING_STEPS.index(s)
            ops.flow_step(1, True, bufs[cur], bufs[nxt], P[:, ci * 128:], plane_strides(P, False), hF[:, ci * 6:], n * 6 * hw,
                          plan.nets_a[ci], plan.pw_inv[s], logdet)
        else:
            ops.flow_step(1, False, bufs[cur], bufs[nxt], None, (0, 0, 0), None, 0, None, plan.pw_inv[s], None)
        if logdet is not None:
            logdet -= (plan.ld_const[s, 0] + plan.ld_const[s, 1]) * float(hw)
        if trace is not None:
            trace.append(bufs[nxt].clone())
        cur, nxt = nxt, (2 if nxt == 1 else 1)       # never write into the caller's tensor
    return bufs[cur], logdet


def encode(plan, gt, ft, conv2d, logdet=None):
    """FlowUpsamplerNet.encode (FlowUpsamplerNet.py:228-274): steps 0 -> 27.  Returns (z, logdet [B])."""
    z = gt.float().contiguous()
    B, _, h, w = z.shape
    hw, n = h * w, len(COUPLING_STEPS)
    if logdet is None:
        logdet = torch.zeros(B, device=z.device, dtype=torch.float32)
    P, hF = precompute(plan, ft, conv2d)
    bufs = [z, torch.empty_like(z), torch.empty_like(z)]
    cur, nxt = 0, 1
    for s in range(N_FLOW_STEPS):
        coupling = s not in NO_COUPLING_STEPS
        logdet += plan.ld_const[s, 0] * float(hw)
        logdet += plan.ld_const[s, 1] * float(hw)
        if coupling:
            ci = COUPLING_STEPS.index(s)
            ops.flow_step(0, True, bufs[cur], bufs[nxt], P[:, ci * 128:], plane_strides(P, False), hF[:, ci * 6:], n * 6 * hw,
                          plan.nets_a[ci], plan.pw_fwd[s], logdet)
        else:
            ops.flow_step(0, False, bufs[cur], bufs[nxt], None, (0, 0, 0), None, 0, None, plan.pw_fwd[s], None)
        cur, nxt = nxt, (2 if nxt == 1 else 1)
    return bufs[cur], logdet


def gaussian_nll(z, mean, logdet):
    """LLFlowVQGAN2.normal_flow objective (LLFlowVQGAN2_arch.py:115-118, flow.py:76-95)."""
    pixels = z.shape[2] * z.shape[3]
    logp = (-0.5 * ((z - mean) ** 2 + math.log(2 * math.pi))).sum(dim=(1, 2, 3))
    return -(logdet + logp) / float(math.log(2.) * pixels)
