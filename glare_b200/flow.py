"""Host side of the conditional flow (FlowUpsamplerNet.py:228-326): parameter packing and the 28-step chains.

Parameter names are the reference's state-dict keys (``flowUpsamplerNet.layers.{s}.actnorm.bias`` ...), so a
``net_G.pth`` loads unchanged.  Everything that depends only on the weights -- exp(logs), the fp64 inverse of
the 1x1 conv (Permutations.py:38), slogdet -- is evaluated ONCE per checkpoint here (batched, on the device) instead
of once per call as the reference does, which also removes the per-step host sync (Permutations.py:25).
"""
import math

import torch

from . import ops

N_FLOW_STEPS = 28
NO_COUPLING_STEPS = (0, 1, 14, 15)          # FlowUpsamplerNet.py:95-106 with LOL.yml K=12, L=2, additionalFlowNoAffine=2
COUPLING_STEPS = tuple(s for s in range(N_FLOW_STEPS) if s not in NO_COUPLING_STEPS)
NET_FLOATS = 9552                           # include/glare_b200.h GLARE_FLOW_NET_FLOATS
HIDDEN = 64


def _stack(sd, keys, device):
    return torch.stack([sd[k].detach().to(device=device, dtype=torch.float32) for k in keys])


def pack_nets(sd, prefixes, has_z, device):
    """The coupling nets `prefixes` (FlowAffineCouplingsAblation.py:143-151) -> packed blocks [n][NET_FLOATS], layout in glare_b200.h.
    Batched over the nets and evaluated on `device`: a handful of tensor ops, no host synchronisation (the stage-2 training step repacks
    after every optimizer step)."""
    n = len(prefixes)
    w1 = _stack(sd, [p + ".0.weight" for p in prefixes], device)                          # [n,64,Cin,3,3]
    w3 = _stack(sd, [p + ".4.weight" for p in prefixes], device)                          # [n,nout,64,3,3]
    nout = w3.shape[1]
    out = torch.zeros((n, NET_FLOATS), dtype=torch.float32, device=device)
    if has_z:
        out[:, 0:576] = w1[:, :, 0].reshape(n, 576)
    out[:, 576:640] = _stack(sd, [p + ".0.actnorm.bias" for p in prefixes], device).reshape(n, HIDDEN)
    out[:, 640:704] = torch.exp(_stack(sd, [p + ".0.actnorm.logs" for p in prefixes], device)).reshape(n, HIDDEN)     # FlowActNorms.py:62
    out[:, 704:4800] = _stack(sd, [p + ".2.weight" for p in prefixes], device)[:, :, :, 0, 0].transpose(1, 2).reshape(n, HIDDEN * HIDDEN)
    out[:, 4800:4864] = _stack(sd, [p + ".2.actnorm.bias" for p in prefixes], device).reshape(n, HIDDEN)
    out[:, 4864:4928] = torch.exp(_stack(sd, [p + ".2.actnorm.logs" for p in prefixes], device)).reshape(n, HIDDEN)
    w3p = torch.zeros((n, HIDDEN, 9, 8), dtype=torch.float32, device=device)
    w3p[:, :, :, :nout] = w3.reshape(n, nout, HIDDEN, 9).permute(0, 2, 3, 1)
    out[:, 4928:9536] = w3p.reshape(n, -1)
    out[:, 9536:9536 + nout] = _stack(sd, [p + ".4.bias" for p in prefixes], device).reshape(n, nout)
    out[:, 9544:9544 + nout] = torch.exp(_stack(sd, [p + ".4.logs" for p in prefixes], device) * 3).reshape(n, nout)  # flow.py:68-70 logscale_factor=3
    return out, w1


def _inv_logdet_3x3(w):
    """w [n,3,3] fp32 -> (inverse [n,3,3] fp32, log|det| [n] fp32), evaluated in fp64 by the cofactor formula with elementwise tensor ops:
    what the reference gets from torch.inverse(w.double()) (Permutations.py:38) and torch.slogdet(w)[1] (:27) to fp32 rounding, without
    the LU library calls (host pointer arrays, error-flag synchronisation) that would break CUDA-graph capture of the training step."""
    m = w.double()
    a, b, c = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    d, e, f = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    g, h, i = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    c00, c01, c02 = e * i - f * h, f * g - d * i, d * h - e * g
    det = a * c00 + b * c01 + c * c02
    adj = torch.stack([torch.stack([c00, c * h - b * i, b * f - c * e], dim=1),
                       torch.stack([c01, a * i - c * g, c * d - a * f], dim=1),
                       torch.stack([c02, b * g - a * h, a * e - b * d], dim=1)], dim=1)
    return (adj / det[:, None, None]).float(), torch.log(det.abs()).float()


class FlowPlan:
    """Device-resident packed parameters of the whole flow (built once per checkpoint; rebuilt per optimizer step in training).  Everything
    is computed on `device` with batched tensor ops -- no `.cpu()` round trips, no host synchronisation."""

    def __init__(self, sd, device, prefix="flowUpsamplerNet"):
        self.prefix = prefix
        self.device = device = torch.device(device)
        pa = ["%s.layers.%d.affine.fAffine" % (prefix, s) for s in COUPLING_STEPS]
        pf = ["%s.layers.%d.affine.fFeatures" % (prefix, s) for s in COUPLING_STEPS]
        self.nets_a, w1a = pack_nets(sd, pa, True, device)
        self.nets_f, w1f = pack_nets(sd, pf, False, device)
        # one dense conv for the ft part of every first layer: rows [ci*128 + 0..63] = NN_A (ft channels of cat([z1, ft]), :137-141),
        # [ci*128 + 64..127] = NN_F
        self.w_pre = torch.cat([w1a[:, :, 1:], w1f], dim=1).reshape(-1, HIDDEN, 3, 3).contiguous()
        # ActNorm + InvertibleConv1x1 of every step -> 16 floats (M row-major, bias, scale), both directions
        steps = range(N_FLOW_STEPS)
        w = _stack(sd, ["%s.layers.%d.invconv.weight" % (prefix, s) for s in steps], device)                     # [28,3,3]
        logs = _stack(sd, ["%s.layers.%d.actnorm.logs" % (prefix, s) for s in steps], device).reshape(N_FLOW_STEPS, 3)
        bias = _stack(sd, ["%s.layers.%d.actnorm.bias" % (prefix, s) for s in steps], device).reshape(N_FLOW_STEPS, 3)
        self.pw_inv = torch.zeros((N_FLOW_STEPS, 16), dtype=torch.float32, device=device)
        self.pw_fwd = torch.zeros((N_FLOW_STEPS, 16), dtype=torch.float32, device=device)
        w_inv, logdet_w = _inv_logdet_3x3(w)
        self.pw_inv[:, 0:9] = w_inv.reshape(N_FLOW_STEPS, 9)                                                     # Permutations.py:38
        self.pw_inv[:, 12:15] = torch.exp(-logs)                                                                 # FlowActNorms.py:64
        self.pw_fwd[:, 0:9] = w.reshape(N_FLOW_STEPS, 9)
        self.pw_fwd[:, 12:15] = torch.exp(logs)
        self.pw_inv[:, 9:12] = bias
        self.pw_fwd[:, 9:12] = bias
        # weight-only logdet terms per step (FlowActNorms.py:66-74, Permutations.py:27,51-53), multiplied by `pixels` at run time
        self.ld_const = torch.stack([logs.sum(dim=1), logdet_w], dim=1)


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
