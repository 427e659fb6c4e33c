"""Torch restatement of the kernel CONTRACTS of csrc/flow_bwd.cu (include/glare_b200.h section 2b), buffer for buffer.  TEST INFRASTRUCTURE:
  * on the CPU it lets tests/test_flow_train_cpu.py run the host orchestration of glare_b200/flow_train.py (offsets, strides, packing,
    gradient assembly) against the specification oracle/flow_backward.py without a GPU;
  * on the GPU it is the per-kernel reference of tests/test_zz_flow_train_gpu.py.
Never imported by the package."""
import torch
import torch.nn.functional as F

from glare_b200.flow import N_FLOW_STEPS, NO_COUPLING_STEPS
from oracle import glare_oracle as O

EPS = 0.0001


def unpack_net(net):
    return {"w1z": net[0:576].view(64, 9), "b1": net[576:640], "s1": net[640:704], "w2t": net[704:4800].view(64, 64), "b2": net[4800:4864],
            "s2": net[4864:4928], "w3": net[4928:9536].view(64, 9, 8), "b3": net[9536:9544], "s3": net[9544:9552]}


def _nchw(x_pc, B, h, w):
    return x_pc.view(B, h, w, -1).permute(0, 3, 1, 2)


def _pc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).reshape(-1, x_nchw.shape[1])


def _scale(hraw):
    return torch.sigmoid(hraw + 2.0) + EPS


class TorchEmuKernels:
    def __init__(self, sd):
        self.sd = sd

    def zeros(self, shape, like):
        return torch.zeros(shape, device=like.device, dtype=torch.float32)

    def empty(self, shape, like):
        return torch.full(shape, float("nan"), device=like.device, dtype=torch.float32)      # unwritten elements must not be consumed

    def encode_chain(self, plan, gt, ft, conv2d):
        B, _, h, w = gt.shape
        zs = torch.empty((N_FLOW_STEPS + 1, B, 3, h, w), device=gt.device)
        zs[0] = gt
        logdet = torch.zeros(B, device=gt.device)
        for s in range(N_FLOW_STEPS):
            zs[s + 1], logdet = O.flow_step_forward(self.sd, "flowUpsamplerNet.layers.%d" % s, zs[s], ft, logdet, s not in NO_COUPLING_STEPS)
        return zs, logdet, conv2d(ft, plan.w_pre)

    def net_fwd(self, pre, pre_off, pre_ld, z1, z1_ld, net, B, h, w, h1, h2, hout):
        n = unpack_net(net)
        a = _nchw(pre.reshape(-1, pre_ld)[:, pre_off:pre_off + 64], B, h, w)
        if z1 is not None:
            a = a + F.conv2d(_nchw(z1.reshape(-1, z1_ld)[:, :1], B, h, w), n["w1z"].view(64, 1, 3, 3), padding=1)
        x1 = torch.relu((a + n["b1"].view(1, -1, 1, 1)) * n["s1"].view(1, -1, 1, 1))
        h1.copy_(_pc(x1))
        x2 = torch.relu((h1 @ n["w2t"] + n["b2"]) * n["s2"])
        h2.copy_(x2)
        w3 = n["w3"].permute(2, 0, 1).reshape(8, 64, 3, 3)
        o = (F.conv2d(_nchw(x2, B, h, w), w3, padding=1) + n["b3"].view(1, -1, 1, 1)) * n["s3"].view(1, -1, 1, 1)
        hout.copy_(_pc(o))

    def point_fwd(self, z_in, pw, hF, B, h, w, t, u, v):
        M, bias, sc = pw[0:9].view(3, 3), pw[9:12], pw[12:15]
        tt = (_pc(z_in) + bias) * sc
        uu = tt @ M.t()
        t.zero_(), u.zero_(), v.zero_()
        t[:, :3], u[:, :3] = tt, uu
        v[:, :3] = uu if hF is None else (uu + hF[:, 0:6:2]) * _scale(hF[:, 1:6:2])

    @staticmethod
    def _affine_bwd(x, shift, hraw, g_y, g_ld):
        sc = _scale(hraw)
        s = sc - EPS
        return g_y * sc, g_y * sc, (g_y * (x + shift) + g_ld / sc) * s * (1.0 - s)

    def coupling_bwd(self, which, g_in, g_z1, x, hraw, g_ld, B, h, w, g_h, g_x):
        g_h.zero_(), g_x.zero_()
        if which == 0:
            go = _pc(g_in)
            g_x[:, 0] = go[:, 0]
            gx, gs, gr = self._affine_bwd(x[:, 1:3], hraw[:, 0:4:2], hraw[:, 1:4:2], go[:, 1:3], g_ld)
            g_x[:, 1:3], g_h[:, 0:4:2], g_h[:, 1:4:2] = gx, gs, gr
        else:
            gy = g_in[:, :3].clone()
            gy[:, 0] += g_z1[:, 0]
            gx, gs, gr = self._affine_bwd(x[:, :3], hraw[:, 0:6:2], hraw[:, 1:6:2], gy, g_ld)
            g_x[:, :3], g_h[:, 0:6:2], g_h[:, 1:6:2] = gx, gs, gr

    def net_bwd(self, g_h, h1, h2, net, B, h, w, g_a3, g_n2, g_a2, g_n1, g_a1, g_pre, pre_off, pre_ld, g_z1):
        n = unpack_net(net)
        g_a3.copy_(g_h * n["s3"])
        w3 = n["w3"].permute(2, 0, 1).reshape(8, 64, 3, 3)
        g_n2.copy_(_pc(F.conv_transpose2d(_nchw(g_a3, B, h, w), w3, padding=1)) * (h2 > 0))
        g_a2.copy_(g_n2 * n["s2"])
        g_n1.copy_((g_a2 @ n["w2t"].t()) * (h1 > 0))
        g_a1.copy_(g_n1 * n["s1"])
        if g_pre is not None:
            g_pre.view(-1, pre_ld)[:, pre_off:pre_off + 64] = g_a1
        if g_z1 is not None:
            g_z1.copy_(_pc(F.conv_transpose2d(_nchw(g_a1, B, h, w), n["w1z"].view(64, 1, 3, 3), padding=1)))

    def point_bwd(self, g_u, t, pw, B, h, w, g_z, sums):
        M, sc = pw[0:9].view(3, 3), pw[12:15]
        gu, tt = g_u[:, :3], t[:, :3]
        gt = gu @ M
        gz = gt * sc
        g_z.copy_(_nchw(gz, B, h, w))
        sums[0:9] += (gu.t() @ tt).reshape(-1)
        sums[9:12] += (gt * tt).sum(0)
        sums[12:15] += gz.sum(0)

    def im2col3x3(self, x, ldx, Cx, B, h, w, col):
        xn = _nchw(x.reshape(-1, ldx)[:, :Cx], B, h, w)
        unf = F.unfold(xn, 3, padding=1).view(B, Cx, 9, h * w)                   # [B][c][t][pixel]
        col.copy_(unf.permute(0, 3, 2, 1).reshape(-1, 9 * Cx))

    def colsum(self, a, lda, b, ldb, Cx, P, out):
        av = a.reshape(-1, lda)[:P, :Cx]
        out[:Cx] += (av if b is None else av * b.reshape(-1, ldb)[:P, :Cx]).sum(0)

    def gemm_tn(self, a, M, b, N, P, out):
        out += a.reshape(P, M).t() @ b.reshape(P, N)
