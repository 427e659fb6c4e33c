"""Stage-2 training of the conditional flow on the library's kernels: the objective of LLFlowVQGAN2.normal_flow
(LLFlowVQGAN2_arch.py:75-122) and its gradients with respect to the latent, the conditioning features, the Gaussian mean and
every flow parameter (reference: torch autograd through FlowUpsamplerNet.encode, FlowUpsamplerNet.py:228-274).

Formulas: oracle/flow_backward.py (CPU specification, checked against autograd and the reference's own gradients).  Kernels:
csrc/flow_bwd.cu, the tensor-core weight gradients of csrc/train_wgrad.cu (skinny / uncovered shapes: the fp32 kernels of flow_bwd.cu and
dcn_bwd.cu) and the tensor-core conv path for the hoisted 64 -> 3072 conv.  This module is the host side: buffer management, the 28-step
loops, and the assembly of the packed gradients into state-dict shaped tensors.

STATUS: the host logic is verified on the CPU against the specification through a torch restatement of every kernel contract and through
the kernel source itself run on the host (tests/flow_train_emu.py, tests/test_flow_train_cpu.py); green on B200 since round 2
(tests/test_zz_flow_train_gpu.py runs tests/flow_train_gpu_check.py in a child process; profiles/r70_train_check.log, 64 checks).
"""
import ctypes
import math

import torch
import torch.nn.functional as F

from . import flow as flowmod
from .flow import COUPLING_STEPS, HIDDEN, N_FLOW_STEPS, NO_COUPLING_STEPS

C = HIDDEN
NPRE = 128 * len(COUPLING_STEPS)            # channels of the hoisted pre-activation tensor: per coupling step 64 (NN_A) + 64 (NN_F)


class CudaKernels:
    """The library's kernels (include/glare_b200.h section 2b).  Pointers are passed as (tensor, element offset) pairs."""

    def __init__(self, mode=4):
        from . import ops
        from ._lib import lib, stream
        self.ops, self.lib, self.stream = ops, lib, stream
        self.mode = mode                 # operand mode of the tensor-core weight gradients: 4 = bf16x3 (fp32-grade), 0 = bf16

    @staticmethod
    def _p(t, off=0):
        return None if t is None else ctypes.c_void_p(t.data_ptr() + 4 * off)

    def _call(self, name, *args):
        self.ops.check(getattr(self.lib(), name)(*args, self.stream()), name)

    def wgrad(self, x_pc, gy_pc, k, B, h, w, chunk=8192):
        """weight gradient [k*k*Ci][N] of a k x k 'same' conv from pixel-major activations x_pc [P][Ci] and output gradients gy_pc [P][N] on
        the tensor cores (csrc/train_wgrad.cu operands + one batched bf16x3 GEMM): the coupling nets' 64 -> 64 1x1 and 64 -> 4 / 6 3x3 layers.
        The 128 x 128-tile split-K fp32 GEMM spent 213 us per call on these skinny shapes, 125 calls per step.  None: shape not covered."""
        P, Ci = x_pc.shape
        N = gy_pc.shape[1]
        if Ci % 32 or k not in (1, 3) or not x_pc.is_contiguous():
            return None
        Np = (N + 31) // 32 * 32
        gy = gy_pc if Np == N and gy_pc.is_contiguous() else F.pad(gy_pc, (0, Np - N)).contiguous()
        G = self.ops.wgrad_conv_tc(x_pc.view(B, h, w, Ci), gy.view(B, h, w, Np), k, 1, k // 2, chunk, mode=self.mode, min_rows=16)
        return None if G is None else G[:, :N]

    def zeros(self, shape, like):
        return torch.zeros(shape, device=like.device, dtype=torch.float32)

    def empty(self, shape, like):
        return torch.empty(shape, device=like.device, dtype=torch.float32)

    def encode_chain(self, plan, gt, ft, conv2d):
        """FlowUpsamplerNet.encode with every step's input kept: -> (zs [29,B,3,h,w], logdet [B], P [B,NPRE,h,w] any dense layout)"""
        z = gt.float().contiguous()
        B, _, h, w = z.shape
        hw, n = h * w, len(COUPLING_STEPS)
        logdet = torch.zeros(B, device=z.device, dtype=torch.float32)
        P, hF = flowmod.precompute(plan, ft, conv2d)
        zs = torch.empty((N_FLOW_STEPS + 1, B, 3, h, w), device=z.device, dtype=torch.float32)
        zs[0].copy_(z)
        for s in range(N_FLOW_STEPS):
            logdet += (plan.ld_const[s, 0] + plan.ld_const[s, 1]) * float(hw)
            if s not in NO_COUPLING_STEPS:
                ci = COUPLING_STEPS.index(s)
                self.ops.flow_step(0, True, zs[s], zs[s + 1], P[:, ci * 128:], flowmod.plane_strides(P, False), hF[:, ci * 6:], n * 6 * hw,
                                   plan.nets_a[ci], plan.pw_fwd[s], logdet)
            else:
                self.ops.flow_step(0, False, zs[s], zs[s + 1], None, (0, 0, 0), None, 0, None, plan.pw_fwd[s], None)
        return zs, logdet, P

    def net_fwd(self, pre, pre_off, pre_ld, z1, z1_ld, net, B, h, w, h1, h2, hout):
        self._call("glare_flow_train_net_fwd_f32", self._p(pre, pre_off), pre_ld, self._p(z1), z1_ld, self._p(net), B, h, w, self._p(h1), self._p(h2),
                   self._p(hout))

    def point_fwd(self, z_in, pw, hF, B, h, w, t, u, v):
        self._call("glare_flow_train_point_fwd_f32", self._p(z_in), self._p(pw), self._p(hF), B, h, w, self._p(t), self._p(u), self._p(v))

    def coupling_bwd(self, which, g_in, g_z1, x, hraw, g_ld, B, h, w, g_h, g_x):
        self._call("glare_flow_train_coupling_bwd_f32", which, self._p(g_in), self._p(g_z1), self._p(x), self._p(hraw), ctypes.c_float(g_ld), B, h, w,
                   self._p(g_h), self._p(g_x))

    def net_bwd(self, g_h, h1, h2, net, B, h, w, g_a3, g_n2, g_a2, g_n1, g_a1, g_pre, pre_off, pre_ld, g_z1):
        self._call("glare_flow_train_net_bwd_f32", self._p(g_h), self._p(h1), self._p(h2), self._p(net), B, h, w, self._p(g_a3), self._p(g_n2),
                   self._p(g_a2), self._p(g_n1), self._p(g_a1), self._p(g_pre, pre_off), pre_ld, self._p(g_z1))

    def point_bwd(self, g_u, t, pw, B, h, w, g_z, sums):
        self._call("glare_flow_train_point_bwd_f32", self._p(g_u), self._p(t), self._p(pw), B, h, w, self._p(g_z), self._p(sums))

    def im2col3x3(self, x, ldx, Cx, B, h, w, col):
        self._call("glare_flow_train_im2col3x3_f32", self._p(x), ldx, Cx, 0, B, h, w, self._p(col))

    def colsum(self, a, lda, b, ldb, Cx, P, out):
        self._call("glare_flow_train_colsum_f32", self._p(a), lda, self._p(b), ldb, Cx, P, self._p(out))

    def gemm_tn(self, a, M, b, N, P, out):
        """out [M][N] += a [P][M]^T b [P][N]: the skinny kernel of csrc/flow_bwd.cu for M <= 32, else the split-K fp32 GEMM of csrc/dcn_bwd.cu"""
        if M <= 32 and N <= 256 and 256 % N == 0:
            self._call("glare_gemm_tn_skinny_f32", self._p(a), self._p(b), P, M, N, self._p(out))
        else:
            self._call("glare_dcnv2_bwd_weight_f32", self._p(a), self._p(b), P, M, N, self._p(out))


class _NetGrads:
    """raw parameter-gradient buffers of all coupling nets, filled net by net by the kernels and converted to the reference's parameter
    layouts ONCE at the end (the per-net views / permutes / clones were ~1 500 tiny torch launches per training step)"""

    def __init__(self, K, like, n_nets):
        self.G3 = K.zeros((n_nets, 9 * C, 8), like)            # Conv2dZeros weight, rows tap * C + c, 8 padded outputs
        self.s8 = K.zeros((n_nets, 2, 8), like)                # its bias / logs sums
        self.G2 = K.zeros((n_nets, C, C), like)                # 1x1 conv weight, [in][out]
        self.s64 = K.zeros((n_nets, 4, C), like)               # the two ActNorms' bias / logs sums
        self.G1 = K.zeros((n_nets, 9, C), like)                # z1 channel of the first conv (nets with a z input only)
        self.keys = [None] * n_nets

    def finish(self, grads):
        G3 = self.G3.view(-1, 9, C, 8).permute(0, 3, 2, 1).reshape(-1, 8, C, 3, 3)
        G2 = self.G2.transpose(1, 2).reshape(-1, C, C, 1, 1)
        G1 = self.G1.transpose(1, 2).reshape(-1, C, 1, 3, 3)
        logs8 = 3.0 * self.s8[:, 1]
        for i, (key, has_z) in enumerate(self.keys):
            nout = 4 if has_z else 6
            grads[key + ".4.weight"] = G3[i, :nout].contiguous()
            grads[key + ".4.bias"] = self.s8[i, 0, :nout]
            grads[key + ".4.logs"] = logs8[i, :nout].view(nout, 1, 1)
            grads[key + ".2.weight"] = G2[i]
            grads[key + ".2.actnorm.bias"] = self.s64[i, 0].view(1, C, 1, 1)
            grads[key + ".2.actnorm.logs"] = self.s64[i, 1].view(1, C, 1, 1)
            grads[key + ".0.actnorm.bias"] = self.s64[i, 2].view(1, C, 1, 1)
            grads[key + ".0.actnorm.logs"] = self.s64[i, 3].view(1, C, 1, 1)
            if has_z:
                grads[key + ".0.weight.z"] = G1[i]


def _net_param_grads(K, acc, i, key, net_has_z, bufs, v, B, h, w, P):
    """parameter gradients of one coupling net from the buffers its data backward left (oracle/flow_backward.py nn_backward), into slot ``i``
    of the stacked buffers ``acc``"""
    h1, h2, hout, g_h, g_a3, g_n2, g_a2, g_n1, g_a1, col576, col9 = bufs
    acc.keys[i] = (key, net_has_z)
    # Conv2dZeros: weight [nout][64][3][3], bias, logs (out = (conv + bias) * exp(3 logs))
    wg = getattr(K, "wgrad", None)               # tensor-core weight gradient where the kernel set has one (the GPU kernels)
    G3 = wg(h2, g_a3, 3, B, h, w) if wg is not None else None
    if G3 is None:
        K.im2col3x3(h2, C, C, B, h, w, col576)
        K.gemm_tn(col576, 9 * C, g_a3, 8, P, acc.G3[i])
    else:
        acc.G3[i].copy_(G3)
    K.colsum(g_a3, 8, None, 0, 8, P, acc.s8[i, 0])
    K.colsum(g_h, 8, hout, 8, 8, P, acc.s8[i, 1])
    # 1x1 conv + ActNorm
    G2 = wg(h1, g_a2, 1, B, h, w) if wg is not None else None
    if G2 is None:
        K.gemm_tn(h1, C, g_a2, C, P, acc.G2[i])
    else:
        acc.G2[i].copy_(G2)
    K.colsum(g_a2, C, None, 0, C, P, acc.s64[i, 0])
    K.colsum(g_n2, C, h2, C, C, P, acc.s64[i, 1])
    K.colsum(g_a1, C, None, 0, C, P, acc.s64[i, 2])
    K.colsum(g_n1, C, h1, C, C, P, acc.s64[i, 3])
    if net_has_z:                               # the z1 input channel of the first 3x3 conv; the ft channels come from the hoisted conv
        K.im2col3x3(v, 4, 1, B, h, w, col9)
        K.gemm_tn(col9, 9, g_a1, C, P, acc.G1[i])


def nll_forward_backward(plan, sd, gt, ft, mean, conv2d, kernels=None, prefix="flowUpsamplerNet"):
    """gt [B,3,h,w] latent, ft [B,64,h,w] conditioning features, mean [B,3,h,w] (color_map) -> (nll [B], z, dL/dgt, dL/dft, dL/dmean,
    {state-dict key: dL/dparam}) for L = nll.mean(); ``conv2d(x, weight)`` is the dense 3x3 'same' conv path (no bias), ``sd`` supplies
    invconv weights for the log|det| term."""
    K = kernels if kernels is not None else CudaKernels()
    B, _, h, w = gt.shape
    hw = h * w
    P = B * hw
    zs, logdet, Ppre = K.encode_chain(plan, gt, ft, conv2d)
    z = zs[N_FLOW_STEPS]
    logp = (-0.5 * ((z - mean) ** 2 + math.log(2 * math.pi))).sum(dim=(1, 2, 3))
    k = 1.0 / (math.log(2.0) * hw)
    nll = -(logdet + logp) * k
    g_ld = -k / B                                                       # dL/dlogdet of every sample
    g_z = ((z - mean) * (k / B)).contiguous()
    g_mean = -g_z

    pre = Ppre.permute(0, 2, 3, 1).contiguous()                          # NHWC [P][NPRE] (a view when the conv path wrote channels_last)
    g_pre = K.empty((B, h, w, NPRE), gt)
    new = lambda c: K.empty((P, c), gt)                                  # noqa: E731
    h1A, h2A, h1F, h2F, g_n2, g_a2, g_n1, g_a1 = (new(C) for _ in range(8))
    hA, hF, g_hA, g_hF, g_a3 = (new(8) for _ in range(5))
    t, u, v, g_v, g_u = (new(4) for _ in range(5))
    g_z1, col576, col9 = new(1), new(9 * C), new(9)
    grads = {}
    acc = _NetGrads(K, gt, 2 * len(COUPLING_STEPS))
    # d log|det W| / dW = W^-T for all 28 invertible 1x1 convs at once: fp64 cofactor inverse (no LU library call: no error-flag
    # synchronisation, CUDA-graph capturable; one batch instead of ~25 tiny tensor ops per step)
    w_all = torch.stack([sd["%s.layers.%d.invconv.weight" % (prefix, s)].to(gt.device, torch.float32) for s in range(N_FLOW_STEPS)])
    w_inv_t = flowmod._inv_logdet_3x3(w_all)[0].transpose(1, 2)
    ld_total = g_ld * B * hw                                            # sum over samples of dL/dlogdet, times the pixel count
    for s in range(N_FLOW_STEPS - 1, -1, -1):
        p = "%s.layers.%d" % (prefix, s)
        pw = plan.pw_fwd[s]
        if s not in NO_COUPLING_STEPS:
            ci = COUPLING_STEPS.index(s)
            netA, netF = plan.nets_a[ci], plan.nets_f[ci]
            offA, offF = ci * 128, ci * 128 + 64
            # forward activations of the step, recomputed from its input
            K.net_fwd(pre, offF, NPRE, None, 0, netF, B, h, w, h1F, h2F, hF)
            K.point_fwd(zs[s], pw, hF, B, h, w, t, u, v)
            K.net_fwd(pre, offA, NPRE, v, 4, netA, B, h, w, h1A, h2A, hA)
            # self coupling -> NN_A -> feature affine -> NN_F
            K.coupling_bwd(0, g_z, None, v, hA, g_ld, B, h, w, g_hA, g_v)
            K.net_bwd(g_hA, h1A, h2A, netA, B, h, w, g_a3, g_n2, g_a2, g_n1, g_a1, g_pre, offA, NPRE, g_z1)
            _net_param_grads(K, acc, 2 * ci, p + ".affine.fAffine", True, (h1A, h2A, hA, g_hA, g_a3, g_n2, g_a2, g_n1, g_a1, col576, col9), v, B, h, w, P)
            K.coupling_bwd(1, g_v, g_z1, u, hF, g_ld, B, h, w, g_hF, g_u)
            K.net_bwd(g_hF, h1F, h2F, netF, B, h, w, g_a3, g_n2, g_a2, g_n1, g_a1, g_pre, offF, NPRE, None)
            _net_param_grads(K, acc, 2 * ci + 1, p + ".affine.fFeatures", False, (h1F, h2F, hF, g_hF, g_a3, g_n2, g_a2, g_n1, g_a1, col576, col9), v, B, h, w, P)
            gu = g_u
        else:
            K.point_fwd(zs[s], pw, None, B, h, w, t, u, v)
            gu = F.pad(g_z.permute(0, 2, 3, 1), (0, 1)).reshape(P, 4).contiguous()
        sums = K.zeros((16,), gt)
        g_z = K.empty((B, 3, h, w), gt)
        K.point_bwd(gu, t, pw, B, h, w, g_z, sums)
        grads[p + ".invconv.weight"] = sums[0:9].view(3, 3) + ld_total * w_inv_t[s]
        grads[p + ".actnorm.logs"] = (sums[9:12] + ld_total).view(1, 3, 1, 1)
        grads[p + ".actnorm.bias"] = sums[12:15].view(1, 3, 1, 1).clone()

    acc.finish(grads)
    # hoisted conv over ft: data gradient on the tensor-core conv path (transpose of a stride-1 'same' conv = the conv with the flipped,
    # transposed filter), weight gradient on the tensor cores (fp32 split-K GEMM over im2col(ft) for kernel sets without them)
    if getattr(plan, "_w_pre_t", None) is None:                        # once per plan: the packed-weight cache of the conv path keys on it
        plan._w_pre_t = plan.w_pre.flip(2, 3).transpose(0, 1).contiguous()
    w_t = plan._w_pre_t
    g_ft = conv2d(g_pre.permute(0, 3, 1, 2), w_t).float()
    ftn = ft.float().permute(0, 2, 3, 1).contiguous()
    wg = getattr(K, "wgrad", None)               # [576][3072] over 25 600 pixels: 90 GFLOP, on the tensor cores where the kernel set has them
    Gp = wg(ftn.view(P, C), g_pre.reshape(P, NPRE), 3, B, h, w) if wg is not None and g_pre.is_contiguous() else None
    if Gp is None:
        K.im2col3x3(ftn, C, C, B, h, w, col576)
        Gp = K.zeros((9 * C, NPRE), gt)
        K.gemm_tn(col576, 9 * C, g_pre, NPRE, P, Gp)
    g_wpre = Gp.view(9, C, NPRE).permute(2, 1, 0).reshape(NPRE, C, 3, 3)
    for ci, s in enumerate(COUPLING_STEPS):
        p = "%s.layers.%d.affine" % (prefix, s)
        grads[p + ".fAffine.0.weight"] = torch.cat([grads.pop(p + ".fAffine.0.weight.z"), g_wpre[ci * 128:ci * 128 + 64]], dim=1).contiguous()
        grads[p + ".fFeatures.0.weight"] = g_wpre[ci * 128 + 64:ci * 128 + 128].contiguous()
    return nll, z, g_z, g_ft, g_mean, grads


class FlowNLL(torch.autograd.Function):
    """``nll = FlowNLL.apply(plan, sd, conv2d, kernels, keys, gt, ft, mean, *params)`` -- the objective of LLFlowVQGAN2.normal_flow as an
    autograd node: the forward runs the library's forward AND backward once (the gradients are linear in the incoming gradient only through
    a per-sample factor, so they are formed for L = nll.mean() and rescaled), the backward hands them out.  ``params`` are the flow
    parameters in the order of ``keys`` (they only mark the graph edges; values are read from ``sd`` / ``plan``).  Gradients of
    ``ft`` / ``mean`` continue into whatever produced them (the reference's ConEncoder1 under torch autograd).

    Restriction: the incoming gradient must be the same for every sample (any mean / sum reduction of the per-sample nll), which is what
    LLFlow_model.optimize_parameters does (LLFlow_model.py:215-232: ``nll.mean()``)."""

    @staticmethod
    def forward(ctx, plan, sd, conv2d, kernels, keys, gt, ft, mean, *params):
        with torch.no_grad():
            nll, z, g_gt, g_ft, g_mean, grads = nll_forward_backward(plan, sd, gt, ft, mean, conv2d, kernels=kernels)
        ctx.batch = gt.shape[0]
        ctx.save_for_backward(g_gt, g_ft, g_mean, *[grads[k] for k in keys])
        return nll

    @staticmethod
    def backward(ctx, g_nll):
        g_gt, g_ft, g_mean, *g_params = ctx.saved_tensors
        if not bool((g_nll == g_nll[0]).all()):
            raise RuntimeError("FlowNLL supports reductions that weigh every sample equally (nll.mean(), nll.sum())")
        scale = g_nll[0] * ctx.batch                                     # stored gradients are those of nll.mean()
        return (None, None, None, None, None, g_gt * scale, g_ft * scale, g_mean * scale) + tuple(g * scale for g in g_params)


def flow_parameter_keys(sd, prefix="flowUpsamplerNet"):
    """state-dict keys of the flow parameters that receive a gradient (the unused ``f`` nets of the noCoupling steps do not)"""
    keys = []
    for s in range(N_FLOW_STEPS):
        p = "%s.layers.%d" % (prefix, s)
        keys += [p + ".actnorm.bias", p + ".actnorm.logs", p + ".invconv.weight"]
        if s not in NO_COUPLING_STEPS:
            for net in ("fAffine", "fFeatures"):
                q = "%s.affine.%s" % (p, net)
                keys += [q + ".0.weight", q + ".0.actnorm.bias", q + ".0.actnorm.logs", q + ".2.weight", q + ".2.actnorm.bias",
                         q + ".2.actnorm.logs", q + ".4.weight", q + ".4.bias", q + ".4.logs"]
    return [k for k in keys if k in sd]
