"""Explicit backward pass of the conditional flow (no autograd graph).  TEST INFRASTRUCTURE -- the CPU specification the
flow-step backward kernels of stage-2 training (BASELINE config 4) are written against.

Forward recurrences are the reference's (FlowStep.py:75-98 normal_flow, FlowActNorms.py:48-100, Permutations.py:21-59,
FlowAffineCouplingsAblation.py:50-151, flow.py:13-70); every gradient formula below is checked against torch autograd of
oracle/glare_oracle.py (tests/test_oracle.py::test_flow_explicit_backward_matches_autograd), which in turn is pinned to the
reference's own autograd (tests/golden/stage2.npz, oracle/gen_golden.py).

Per step the kernels need: the step input z_in, the conditioning features ft, and from the forward pass the pre-ReLU
activations n1, n2 of both coupling nets and their outputs h (4 / 6 channels) -- at the stage-2 crop (latent 80x80, batch 4)
that is 24 steps x 2 nets x 128 channels x 25 600 pixels x 4 B = 630 MB, kept instead of recomputed.
"""
import math

import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

from .glare_oracle import N_FLOW_STEPS, NO_COUPLING_STEPS

EPS = 0.0001                    # affine_eps, FlowAffineCouplingsAblation.py:31


# ------------------------------------------------------------------------------------------ coupling nets
def nn_forward(sd, p, x):
    """CondAffineSeparatedAndCond.F with the activations its backward needs (FlowAffineCouplingsAblation.py:143-151)"""
    e1, e2 = torch.exp(sd[p + ".0.actnorm.logs"]), torch.exp(sd[p + ".2.actnorm.logs"])
    n1 = (F.conv2d(x, sd[p + ".0.weight"], None, padding=1) + sd[p + ".0.actnorm.bias"]) * e1
    h1 = torch.relu(n1)
    n2 = (F.conv2d(h1, sd[p + ".2.weight"], None, padding=0) + sd[p + ".2.actnorm.bias"]) * e2
    h2 = torch.relu(n2)
    out = (F.conv2d(h2, sd[p + ".4.weight"], sd[p + ".4.bias"], padding=1)) * torch.exp(sd[p + ".4.logs"] * 3)
    return out, (x, n1, h1, n2, h2, out)


def nn_backward(sd, p, cache, g_out, grads):
    """g_out = dL/d(net output) -> dL/d(net input); parameter gradients accumulated into `grads` under the state-dict keys"""
    x, n1, h1, n2, h2, out = cache
    w1, w2, w3 = sd[p + ".0.weight"], sd[p + ".2.weight"], sd[p + ".4.weight"]
    e1, e2, e3 = torch.exp(sd[p + ".0.actnorm.logs"]), torch.exp(sd[p + ".2.actnorm.logs"]), torch.exp(sd[p + ".4.logs"] * 3)
    red = (0, 2, 3)
    # Conv2dZeros: out = (conv(h2) + bias) * exp(3 logs)                                          flow.py:55-70
    _acc(grads, p + ".4.logs", 3.0 * (g_out * out).sum(red).view_as(sd[p + ".4.logs"]))
    g_a3 = g_out * e3
    _acc(grads, p + ".4.bias", g_a3.sum(red))
    _acc(grads, p + ".4.weight", conv2d_weight(h2, w3.shape, g_a3, padding=1))
    g_n2 = conv2d_input(h2.shape, w3, g_a3, padding=1) * (n2 > 0)
    # ActNorm (x + bias) * exp(logs): d/dlogs = n, d/dbias = exp(logs)                            FlowActNorms.py:48-72
    _acc(grads, p + ".2.actnorm.logs", (g_n2 * n2).sum(red).view_as(sd[p + ".2.actnorm.logs"]))
    g_a2 = g_n2 * e2
    _acc(grads, p + ".2.actnorm.bias", g_a2.sum(red).view_as(sd[p + ".2.actnorm.bias"]))
    _acc(grads, p + ".2.weight", conv2d_weight(h1, w2.shape, g_a2, padding=0))
    g_n1 = conv2d_input(h1.shape, w2, g_a2, padding=0) * (n1 > 0)
    _acc(grads, p + ".0.actnorm.logs", (g_n1 * n1).sum(red).view_as(sd[p + ".0.actnorm.logs"]))
    g_a1 = g_n1 * e1
    _acc(grads, p + ".0.actnorm.bias", g_a1.sum(red).view_as(sd[p + ".0.actnorm.bias"]))
    _acc(grads, p + ".0.weight", conv2d_weight(x, w1.shape, g_a1, padding=1))
    return conv2d_input(x.shape, w1, g_a1, padding=1)


def _acc(grads, key, g):
    grads[key] = grads[key] + g if key in grads else g


def _affine_forward(x, h):
    """(x + shift) * scale with scale = sigmoid(h[1::2] + 2) + eps, shift = h[0::2]               :121-135"""
    scale, shift = torch.sigmoid(h[:, 1::2] + 2.0) + EPS, h[:, 0::2]
    return (x + shift) * scale, scale, shift


def _affine_backward(x, scale, shift, g_y, g_ld):
    """-> (g_x, g_h): g_ld [B] is dL/dlogdet of the sample (logdet += sum log scale)"""
    g_scale = g_y * (x + shift) + g_ld.view(-1, 1, 1, 1) / scale
    s = scale - EPS
    g_h = torch.empty((x.shape[0], 2 * x.shape[1]) + tuple(x.shape[2:]), dtype=x.dtype)
    g_h[:, 0::2] = g_y * scale                      # shift
    g_h[:, 1::2] = g_scale * s * (1.0 - s)          # through the sigmoid
    return g_y * scale, g_h


# ------------------------------------------------------------------------------------------ one FlowStep
def step_forward(sd, p, z, ft, coupling):
    """FlowStep.normal_flow (FlowStep.py:75-98) -> (z_out, logdet increment [B], cache)"""
    pixels = z.shape[2] * z.shape[3]
    logs, w = sd[p + ".actnorm.logs"], sd[p + ".invconv.weight"]
    t = (z + sd[p + ".actnorm.bias"]) * torch.exp(logs)
    u = F.conv2d(t, w.view(3, 3, 1, 1))
    ld = (logs.sum() + torch.slogdet(w)[1]) * pixels * torch.ones(z.shape[0])
    cache = {"t": t, "u": u, "pixels": pixels}
    if not coupling:
        return u, ld, cache
    hF, cF = nn_forward(sd, p + ".affine.fFeatures", ft)
    v, sF, bF = _affine_forward(u, hF)
    hA, cA = nn_forward(sd, p + ".affine.fAffine", torch.cat([v[:, :1], ft], dim=1))
    y2, sA, bA = _affine_forward(v[:, 1:], hA)
    ld = ld + torch.log(sF).sum(dim=(1, 2, 3)) + torch.log(sA).sum(dim=(1, 2, 3))
    cache.update(cF=cF, cA=cA, v=v, sF=sF, bF=bF, sA=sA, bA=bA)
    return torch.cat([v[:, :1], y2], dim=1), ld, cache


def step_backward(sd, p, cache, coupling, g_out, g_ld, grads):
    """g_out = dL/dz_out, g_ld [B] = dL/dlogdet -> (dL/dz_in, dL/dft or None)"""
    logs, w = sd[p + ".actnorm.logs"], sd[p + ".invconv.weight"]
    g_ft = None
    g_u = g_out
    if coupling:
        v = cache["v"]
        g_y2, g_hA = _affine_backward(v[:, 1:], cache["sA"], cache["bA"], g_out[:, 1:], g_ld)
        g_xA = nn_backward(sd, p + ".affine.fAffine", cache["cA"], g_hA, grads)          # input was cat([z1, ft])
        g_v = torch.cat([g_out[:, :1] + g_xA[:, :1], g_y2], dim=1)
        g_u, g_hF = _affine_backward(cache["u"], cache["sF"], cache["bF"], g_v, g_ld)
        g_ft = g_xA[:, 1:] + nn_backward(sd, p + ".affine.fFeatures", cache["cF"], g_hF, grads)
    # InvertibleConv1x1: u = W t, logdet += pixels * log|det W|  (d log|det W| / dW = W^-T)       Permutations.py:21-59
    t = cache["t"]
    g_w = torch.einsum("bohw,bihw->oi", g_u, t) + g_ld.sum() * cache["pixels"] * torch.inverse(w.double()).t().float()
    _acc(grads, p + ".invconv.weight", g_w)
    g_t = F.conv2d(g_u, w.t().contiguous().view(3, 3, 1, 1))
    # ActNorm2d: t = (z + bias) * exp(logs), logdet += pixels * sum(logs)                          FlowActNorms.py:48-100
    _acc(grads, p + ".actnorm.logs", ((g_t * t).sum((0, 2, 3)) + g_ld.sum() * cache["pixels"]).view_as(logs))
    g_z = g_t * torch.exp(logs)
    _acc(grads, p + ".actnorm.bias", g_z.sum((0, 2, 3)).view_as(sd[p + ".actnorm.bias"]))
    return g_z, g_ft


# ------------------------------------------------------------------------------------------ the chain and its objective
def nll_forward_backward(sd, gt, ft, mean, p="flowUpsamplerNet"):
    """mean over the batch of LLFlowVQGAN2.normal_flow's objective (LLFlowVQGAN2_arch.py:115-118, flow.py:76-95) and its
    gradients: -> (nll [B], z, dL/dgt, dL/dft, dL/dmean, {state-dict key: dL/dparam}) with L = nll.mean()"""
    B, _, h, w = gt.shape
    pixels = h * w
    z, logdet, caches = gt, torch.zeros(B), []
    for s in range(N_FLOW_STEPS):
        z, ld, c = step_forward(sd, "%s.layers.%d" % (p, s), z, ft, s not in NO_COUPLING_STEPS)
        logdet = logdet + ld
        caches.append(c)
    logp = (-0.5 * ((z - mean) ** 2 + math.log(2 * math.pi))).sum(dim=(1, 2, 3))
    k = 1.0 / (math.log(2.0) * pixels)
    nll = -(logdet + logp) * k
    g_ld = torch.full((B,), -k / B)                                   # dL/dlogdet
    g_z = (z - mean) * (k / B)                                        # dL/dz = -k/B * dlogp/dz
    g_mean = -g_z
    grads, g_ft = {}, torch.zeros_like(ft)
    for s in range(N_FLOW_STEPS - 1, -1, -1):
        g_z, g = step_backward(sd, "%s.layers.%d" % (p, s), caches[s], s not in NO_COUPLING_STEPS, g_z, g_ld, grads)
        if g is not None:
            g_ft = g_ft + g
    return nll, z, g_z, g_ft, g_mean, grads
