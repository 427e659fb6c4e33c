"""Torch restatement of the kernel-level primitives ("leaves") of glare_b200/encoder_train.py.  TEST INFRASTRUCTURE (see flow_train_emu.py)."""
import torch
import torch.nn.functional as F


def _tv_offsets(offset, dg):
    """the reference's offset layout (o1 | o2 halves, deformableDecoder_arch.py:144-146 + deform_conv_cuda_kernel.cu:596-601: per group g the
    channels [g*18 + 2*t, g*18 + 2*t + 1] = (dy, dx) of tap t) is already torchvision's: identity, kept as the single place that says so"""
    return offset


class TorchLeaves:
    def conv_same(self, x, w, b=None):
        return F.conv2d(x, w, b, padding=w.shape[2] // 2)

    def conv_down(self, x, w, b=None):
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)

    def attention(self, q, k, v):
        B, C, h, w = q.shape
        s = torch.softmax(torch.bmm(q.reshape(B, C, h * w).permute(0, 2, 1), k.reshape(B, C, h * w)) * (int(C) ** (-0.5)), dim=2)
        return torch.bmm(v.reshape(B, C, h * w), s.permute(0, 2, 1)).reshape(B, C, h, w)

    def colsum(self, x):
        return x.sum(dim=0)

    def relu(self, x):
        return F.relu(x)

    def relu_bwd(self, y, gy):
        return gy * (y > 0)

    def maxpool2(self, x):
        y, where = F.max_pool2d(x, 2, 2, return_indices=True)
        return y, (where, x.shape)

    def maxpool2_bwd(self, gy, saved):
        return F.max_unpool2d(gy, saved[0], 2, 2, output_size=saved[1][2:])

    def avgpool2(self, x):
        return F.avg_pool2d(x, (2, 2))

    def up2(self, x):
        return F.interpolate(x, scale_factor=2.0, mode="nearest")

    def up2_adjoint(self, gy):
        B, C, H2, W2 = gy.shape
        return gy.reshape(B, C, H2 // 2, 2, W2 // 2, 2).sum(dim=(3, 5))

    def dcn_fwd(self, x, offmask_raw, w, b, dg):
        from torchvision.ops import deform_conv2d
        n_off = offmask_raw.shape[1] // 3 * 2
        return deform_conv2d(x, _tv_offsets(offmask_raw[:, :n_off], dg), w, b, padding=1, mask=torch.sigmoid(offmask_raw[:, n_off:]))

    def dcn_bwd(self, x, offset, mask, w, gy, dg):
        from torchvision.ops import deform_conv2d
        with torch.enable_grad():
            leaves = [t.detach().clone().requires_grad_(True) for t in (x, offset, mask, w)]
            b = torch.zeros(w.shape[0], requires_grad=True)
            y = deform_conv2d(leaves[0], _tv_offsets(leaves[1], dg), leaves[3], b, padding=1, mask=leaves[2])
            gi, go, gm, gw, gb = torch.autograd.grad(y, leaves[:3] + [leaves[3], b], gy)
        return gi, go, gm, gw, gb

    def gemm_tn(self, a, b):
        return a.t() @ b

    def gemm_nt(self, a, b, rows_hw):
        return a @ b.t()

    def gemm_nt_ta(self, at, b, rows_hw):
        return at.t() @ b.t()

    def im2col(self, x_nhwc, k, stride, pad, Ho, Wo):
        B, H, W, C = x_nhwc.shape
        need_h, need_w = (Ho - 1) * stride + k - pad - H, (Wo - 1) * stride + k - pad - W
        xp = F.pad(x_nhwc.permute(0, 3, 1, 2), (pad, max(need_w, 0), pad, max(need_h, 0)))
        unf = F.unfold(xp, k, stride=stride).view(B, C, k * k, Ho * Wo)           # [B][c][t][pixel]
        return unf.permute(0, 3, 2, 1).reshape(B * Ho * Wo, k * k * C).contiguous()

    def gn_fwd(self, x_nhwc, gamma, beta, swish):
        B, H, W, C = x_nhwc.shape
        xg = x_nhwc.double().reshape(B, H * W, 32, C // 32)
        stats = torch.stack([xg.sum(dim=(1, 3)), (xg * xg).sum(dim=(1, 3))], dim=-1)          # [B][32][2] fp64: sum, sum of squares
        y = F.group_norm(x_nhwc.permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-6)
        if swish:
            y = y * torch.sigmoid(y)
        return y.permute(0, 2, 3, 1).contiguous(), stats

    def gn_bwd(self, x_nhwc, gy_nhwc, stats, gamma, beta, swish):
        B, H, W, C = x_nhwc.shape
        cnt = float(H * W * (C // 32))
        mean = (stats[..., 0] / cnt)
        rstd = 1.0 / torch.sqrt((stats[..., 1] / cnt - mean * mean).clamp_min(0) + 1e-6)
        mean_c = mean.float().repeat_interleave(C // 32, dim=1).view(B, 1, 1, C)
        rstd_c = rstd.float().repeat_interleave(C // 32, dim=1).view(B, 1, 1, C)
        xh = (x_nhwc - mean_c) * rstd_c
        n = xh * gamma + beta
        if swish:
            s = torch.sigmoid(n)
            dn = gy_nhwc * (s * (1 + n * (1 - s)))
        else:
            dn = gy_nhwc
        dxh = dn * gamma
        g = lambda t: t.reshape(B, H * W, 32, C // 32).mean(dim=(1, 3)).repeat_interleave(C // 32, dim=1).view(B, 1, 1, C)    # noqa: E731
        gx = rstd_c * (dxh - g(dxh) - xh * g(dxh * xh))
        return gx, (dn * xh).sum(dim=(0, 1, 2)), dn.sum(dim=(0, 1, 2))

    def softmax_rows(self, S, scale):
        return torch.softmax(S * scale, dim=1)

    def softmax_bwd(self, P, dP, scale):
        return scale * P * (dP - (dP * P).sum(dim=1, keepdim=True))
