"""Stage-2 training step on the library's kernels: condition encoder (ConEncoder1, ConditionEncoder.py:46-55 over the taming Encoder,
encoder_decoder.py:406-442) forward with a tape, the flow objective through glare_b200/flow_train.py, and the backward pass of every block
-- what the reference gets from torch autograd in LLFlow_model.optimize_parameters (LLFlow_model.py:215-232) around
LLFlowVQGAN2.normal_flow (LLFlowVQGAN2_arch.py:75-122).

Layering: ``Leaves`` are the kernel-level primitives (tensor-core conv path, split-K GEMM, GroupNorm / softmax forward and backward kernels);
the block-level backward rules above them (how a conv's data gradient becomes a conv with the flipped, transposed filter, how a stride-2
Downsample's becomes a conv over the zero-interleaved gradient, the five matmuls of the attention backward ...) are plain Python shared by
every backend, so the CPU test (tests/test_encoder_train_cpu.py) runs exactly this logic with torch / host-compiled leaves against torch
autograd of the oracle and the reference's own gradients.

STATUS: green on B200 (tests/test_zz_flow_train_gpu.py, profiles/r70_train_check.log); 94 ms per objective + gradient evaluation at BASELINE
config 4 (batch 4 x 320x320) with fp32-grade operands, 78 ms with bf16 operands (profiles/r71_train_probe*_kernel_breakdown.txt).
"""
import ctypes
import os
import weakref

import torch
import torch.nn.functional as F

from . import flow_train


def _nhwc(x):
    """logical [B,C,H,W] -> contiguous [B,H,W,C] fp32"""
    return x.float().permute(0, 2, 3, 1).contiguous()


def _nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2)


_WT = {}
_NO_CACHE = False          # set while a training step is being captured into a CUDA graph: every replay recomputes from the current weights


def _flipped_transposed(w):
    """[Co,Ci,k,k] -> [Ci,Co,k,k] with the taps reversed; one cached copy per weight (storage address, shape), REPLACED when the optimizer's
    in-place update moves the tensor's version -- keying on the version kept a flipped copy (and, through it, a packed operand copy in the
    conv path's cache) of every weight per training step: 0.18 GB per stage-2 step until the table was cleared"""
    if _NO_CACHE:
        return w.flip(2, 3).transpose(0, 1).contiguous()
    key = (w.data_ptr(), tuple(w.shape))
    ent = _WT.get(key)
    if ent is None or ent[0]() is not w or ent[1] != w._version:
        if ent is None and len(_WT) > 1024:
            for k in [k for k, e in _WT.items() if e[0]() is None]:
                del _WT[k]
        ent = _WT[key] = (weakref.ref(w), w._version, w.flip(2, 3).transpose(0, 1).contiguous())
    return ent[2]


class CudaLeaves:
    """kernel-level primitives on the GPU (libglare_b200.so through glare_b200.ops / the dense backend)"""

    def __init__(self, dense):
        from . import ops
        from ._lib import lib, stream
        self.dense, self.ops, self.lib, self.stream = dense, ops, lib, stream
        import os
        # weight gradients on the tensor cores (gemm_tn_tc) unless GLARE_WGRAD_FMA=1 selects the fp32 split-K GEMM (the correctness baseline):
        # 209 vs 228 ms per stage-2 step at batch 4 x 320x320 (profiles/r46_train_probe*.txt)
        # (mode 4 = fp32-grade bf16x3 operands, the default; mode 0 = bf16 operands, the bf16 training configuration of BASELINE config 4)
        self.wgrad_tc = not os.environ.get("GLARE_WGRAD_FMA") and getattr(dense, "mode", None) in (0, 4)

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _call(self, name, *args):
        self.ops.check(getattr(self.lib(), name)(*args, self.stream()), name)

    # dense
    def conv_same(self, x, w, b=None):
        return self.dense.conv2d(x, w, b, stride=1, padding=w.shape[2] // 2, gn_stats=False).float()   # the tape's Normalize takes its own statistics

    def conv_down(self, x, w, b=None):
        return self.dense.downsample_conv(x, w, b, gn_stats=False).float()          # shapes outside the kernel's coverage raise (no library path)

    def attention(self, q, k, v):
        # exact softmax: the fused-softmax fast path relies on a device flag the caller must read and act on (engine.infer does); a training
        # forward cannot recompute after the fact, and the backward (Tape._attn_bwd) recomputes P with the exact softmax anyway
        return self.dense.attention(q, k, v, fused=False).float()

    def gemm_tn(self, a, b):
        """a [P][M], b [P][N] -> a^T b [M][N] (fp32 split-K GEMM, csrc/dcn_bwd.cu)"""
        P, M = a.shape
        if self.wgrad_tc and M >= 128 and b.shape[1] >= 32:
            return self.gemm_tn_tc(a, b)
        N = b.shape[1]
        out = torch.zeros((M, N), device=a.device, dtype=torch.float32)
        if M <= 32 and N <= 256 and 256 % N == 0 and a.is_contiguous() and b.is_contiguous():     # conv_in: 27 im2col columns of a 3-channel input
            self._call("glare_gemm_tn_skinny_f32", self._p(a), self._p(b), P, M, N, self._p(out))
        else:
            self._call("glare_dcnv2_bwd_weight_f32", self._p(a), self._p(b), P, M, N, self._p(out))
        return out

    def colsum(self, x):
        """x [P][C] -> column sums [C] (bias gradients): csrc/train_enc.cu colsum_kernel, one launch per tensor; widths that are not a multiple
        of 4 (3-channel heads) are tiny reductions left to torch.  (Round 1 ran these through the 128 x 128-tile split-K GEMM with a ones
        vector: 35 ms of a 194 ms step; round 2 first used the flow path's 64-column kernel, C / 64 launches per tensor.)"""
        P, C = x.shape
        if C % 4 or not x.is_contiguous():
            return x.sum(dim=0)
        out = torch.zeros((C,), device=x.device, dtype=torch.float32)
        self._call("glare_colsum_f32", self._p(x), P, C, self._p(out))
        return out

    def gemm_tn_tc(self, a, b, chunk=8192):
        """a [P][M], b [P][N] -> a^T b [M][N] on the tensor cores: the reduction over the P pixels is cut into chunks that become the BATCH of one
        tcgen05 GEMM launch with per-sample weights (sample c: rows = a_c^T [M][chunk], weights = b_c^T [N][chunk]), the per-chunk products are
        summed in fp32.  Default in the fp32-grade (bf16x3) mode; GLARE_WGRAD_FMA=1 selects the fp32 split-K GEMM instead."""
        P, M = a.shape
        N = b.shape[1]
        chunk = max(32, min(chunk, (P + 31) // 32 * 32) // 32 * 32)
        nch = (P + chunk - 1) // chunk
        pad = nch * chunk - P
        if pad:
            a, b = F.pad(a, (0, 0, 0, pad)), F.pad(b, (0, 0, 0, pad))
        at = a.view(nch, chunk, M).transpose(1, 2).contiguous()             # [nch][M][chunk]
        bt = b.view(nch, chunk, N).transpose(1, 2).contiguous()             # [nch][N][chunk]
        mode = self.dense.mode
        a_hi, a_lo = self.ops.conv_prep_act(mode, at)
        b_hi, b_lo = self.ops.conv_prep_act(mode, bt)
        rows_w = 16 if M % 16 == 0 else 1
        ldy = (N + 3) // 4 * 4
        y = torch.empty((nch, M, ldy), device=a.device, dtype=torch.float32)
        self.ops.conv2d_nhwc_tc_ex(mode, a_hi, a_lo, b_hi, b_lo, y, nch, M // rows_w, rows_w, chunk, N, ldy, N * chunk)
        return y[:, :, :N].sum(dim=0)

    def wgrad_conv(self, x_nhwc, gy_nhwc, k, stride, pad, chunk=8192):
        """weight gradient of a k x k conv, [k*k*Ci][Co] (tap-major rows), on the tensor cores without fp32 im2col columns (ops.wgrad_conv_tc).
        None when the shape is outside that path (the caller uses im2col + gemm_tn)."""
        if not self.wgrad_tc:
            return None
        return self.ops.wgrad_conv_tc(x_nhwc, gy_nhwc, k, stride, pad, chunk, mode=self.dense.mode)

    def gemm_nt(self, a, b, rows_hw):
        """a [R][K], b [N][K] -> a b^T [R][N] on the tcgen05 GEMM path (R = rows_hw[0] * rows_hw[1], the tile walk needs the 2-D factorisation)"""
        R, K = a.shape
        N = b.shape[0]
        padk = (-K) % 32
        if padk:
            a, b = F.pad(a, (0, padk)), F.pad(b, (0, padk))
        mode = self.dense.mode
        a_hi, a_lo = self.ops.conv_prep_act(mode, a.contiguous())
        b_hi, b_lo = self.ops.conv_prep_act(mode, b.contiguous())
        ldy = (N + 3) // 4 * 4
        y = torch.empty((R, ldy), device=a.device, dtype=torch.float32)
        self.ops.conv2d_nhwc_tc_ex(mode, a_hi, a_lo, b_hi, b_lo, y, 1, rows_hw[0], rows_hw[1], K + padk, N, ldy, 0)
        return y[:, :N]

    def gemm_nt_ta(self, at, b, rows_hw):
        """at [K][R], b [N][K] -> at^T b^T [R][N]: the left operand is given as its TRANSPOSE and its tensor-core operand is written directly
        from that layout (csrc/train_wgrad.cu with k = 1: pixels = the K index), instead of a transposed fp32 copy followed by the operand
        conversion -- the two N x N transposes per sample of the attention backward were 4.3 ms of the stage-2 step."""
        K, R = at.shape
        N = b.shape[0]
        mode = self.dense.mode
        if mode not in (0, 4) or R % 32 or not at.is_contiguous() or os.environ.get("GLARE_NO_GEMM_TA"):     # (the env switch: A/B timing)
            return self.gemm_nt(at.t().contiguous(), b, rows_hw)
        q, e2 = (32, 2) if mode == 4 else (64, 1)
        chunk = (K + q - 1) // q * q
        a_op = torch.empty((1, R, e2 * chunk), device=at.device, dtype=torch.bfloat16)
        self._call("glare_im2col_t_operand_bf16x3" if mode == 4 else "glare_im2col_t_operand_bf16", self._p(at), 1, 1, K, R, 1, 1, 0, 1, K, chunk,
                   self._p(a_op))
        bp = F.pad(b, (0, chunk - K)) if chunk != K else b
        b_hi, b_lo = self.ops.conv_prep_act(mode, bp.contiguous())
        ldy = (N + 3) // 4 * 4
        y = torch.empty((R, ldy), device=at.device, dtype=torch.float32)
        self.ops.conv2d_nhwc_tc_ex(mode, a_op, None, b_hi, b_lo, y, 1, rows_hw[0], rows_hw[1], chunk, N, ldy, 0)
        return y[:, :N]

    # modulated deformable convolution (stage 3, decoder_train.py)
    def dcn_fwd(self, x, offmask_raw, w, b, dg):
        """DCNv2Pack after conv_offset: the tensor-core kernel that takes the raw conv_offset output (offsets | mask logits), else the fp32
        operator on (offset, sigmoid(mask))"""
        y = self.dense.dcn_pack(x, offmask_raw, w, b, dg) if hasattr(self.dense, "dcn_pack") else None
        if y is not None:
            return y.float()
        n_off = offmask_raw.shape[1] // 3 * 2
        return self.ops.modulated_deform_conv(x.contiguous(), offmask_raw[:, :n_off].contiguous(), torch.sigmoid(offmask_raw[:, n_off:]).contiguous(),
                                              w, b, 1, 1, 1, 1, dg)

    def dcn_bwd(self, x, offset, mask, w, gy, dg):
        """(grad_input, grad_offset, grad_mask, grad_weight, grad_bias): csrc/dcn_bwd.cu through dcn_backward.py"""
        from .dcn_backward import dcn_backward
        return dcn_backward(x, offset, mask, w, gy, dg, True)

    # elementwise / pooling pieces of the stage-3 step (csrc/loss.cu); logical NCHW in and out, NHWC underneath
    def relu(self, x):
        xn = _nhwc(x)
        y = torch.empty_like(xn)
        self._call("glare_relu_f32", self._p(xn), None, xn.numel(), self._p(y))
        return _nchw(y)

    def relu_bwd(self, y, gy):
        yn, gn = _nhwc(y), _nhwc(gy)
        gx = torch.empty_like(yn)
        self._call("glare_relu_f32", self._p(yn), self._p(gn), yn.numel(), self._p(gx))
        return _nchw(gx)

    def maxpool2(self, x):
        xn = _nhwc(x)
        B, H, W, C = xn.shape
        y = torch.empty((B, H // 2, W // 2, C), device=xn.device, dtype=torch.float32)
        idx = torch.empty((B, H // 2, W // 2, C), device=xn.device, dtype=torch.uint8)
        self._call("glare_maxpool2_nhwc_f32", self._p(xn), B, H, W, C, self._p(y), self._p(idx))
        return _nchw(y), (idx, (B, H, W, C))

    def maxpool2_bwd(self, gy, saved):
        idx, (B, H, W, C) = saved
        gn = _nhwc(gy)                                                         # held until the call is issued
        gx = torch.empty((B, H, W, C), device=gy.device, dtype=torch.float32)
        self._call("glare_maxpool2_nhwc_bwd_f32", self._p(gn), self._p(idx), B, H, W, C, self._p(gx))
        return _nchw(gx)

    def avgpool2(self, x):
        x = x.float().contiguous()                                            # [B][C][H][W] planes
        B, C, H, W = x.shape
        y = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=torch.float32)
        self._call("glare_avgpool2_f32", self._p(x), B * C, H, W, self._p(y))
        return y

    def up2(self, x):
        xn = _nhwc(x)
        B, H, W, C = xn.shape
        y = torch.empty((B, 2 * H, 2 * W, C), device=xn.device, dtype=torch.float32)
        self._call("glare_up2_nhwc_f32", self._p(xn), B, H, W, C, 0, self._p(y))
        return _nchw(y)

    def up2_adjoint(self, gy):
        gn = _nhwc(gy)
        B, H2, W2, C = gn.shape
        gx = torch.empty((B, H2 // 2, W2 // 2, C), device=gn.device, dtype=torch.float32)
        self._call("glare_up2_nhwc_f32", self._p(gn), B, H2 // 2, W2 // 2, C, 1, self._p(gx))
        return _nchw(gx)

    # memory-bound kernels
    def im2col(self, x_nhwc, k, stride, pad, Ho, Wo):
        B, H, W, C = x_nhwc.shape
        col = torch.empty((B * Ho * Wo, k * k * C), device=x_nhwc.device, dtype=torch.float32)
        self._call("glare_im2col_nhwc_f32", self._p(x_nhwc), B, H, W, C, k, stride, pad, Ho, Wo, self._p(col))
        return col

    def gn_fwd(self, x_nhwc, gamma, beta, swish):
        B, H, W, C = x_nhwc.shape
        stats = self.ops.gn_stats(x_nhwc, B, H * W, C)
        y, _ = self.ops.gn_apply(1, x_nhwc, stats, gamma, beta, swish, B, H * W, C)
        return y, stats

    def gn_bwd(self, x_nhwc, gy_nhwc, stats, gamma, beta, swish):
        B, H, W, C = x_nhwc.shape
        sums = torch.empty((B, 32, 2), device=x_nhwc.device, dtype=torch.float64)
        gx = torch.empty_like(x_nhwc)
        dg, db = torch.zeros_like(gamma), torch.zeros_like(beta)
        self._call("glare_gn_bwd_nhwc_f32", self._p(x_nhwc), self._p(gy_nhwc), self._p(stats), self._p(gamma), self._p(beta), ctypes.c_float(1e-6),
                   1 if swish else 0, B, H * W, C, 32, self._p(sums), self._p(gx), self._p(dg), self._p(db))
        return gx, dg, db

    def softmax_rows(self, S, scale):
        R, N = S.shape
        ld = (N + 3) // 4 * 4                                              # the kernel reads / writes rows of a multiple of 4 floats
        Sp = F.pad(S, (0, ld - N)).contiguous() if ld != N else S.contiguous()
        P = torch.empty_like(Sp)
        self.ops.attn_softmax_rows(1, Sp, R, ld, N, ld, scale, P, None, ld)
        return P[:, :N].contiguous() if ld != N else P

    def softmax_bwd(self, P, dP, scale):
        dS = torch.empty_like(P)
        self._call("glare_attn_softmax_bwd_f32", self._p(P), self._p(dP), P.shape[0], P.shape[1], P.shape[1], ctypes.c_float(scale), self._p(dS))
        return dS


# ---------------------------------------------------------------------------------------------------------------- tape
class Tape:
    """reverse-mode bookkeeping for the encoder's five block types; gradients of parameters are accumulated under their state-dict keys"""

    def __init__(self, leaves, sd):
        self.L, self.sd, self.ops, self.grads = leaves, sd, [], {}

    def _acc(self, key, g):
        self.grads[key] = self.grads[key] + g if key in self.grads else g

    # -- convolution (encoder_decoder.py nn.Conv2d 3x3 / 1x1 stride 1, and Downsample :68-72) -----------------------------------------
    def conv(self, p, x, down=False, need_gx=True, need_gw=True):
        w, b = self.sd[p + ".weight"], self.sd.get(p + ".bias")
        y = self.L.conv_down(x, w, b) if down else self.L.conv_same(x, w, b)
        self.ops.append(("conv", p, x, y, down, need_gx, need_gw))
        return y

    def _conv_bwd(self, p, x, y, down, need_gx, gy, need_gw=True):
        w = self.sd[p + ".weight"]
        Co, Ci, k, _ = w.shape
        B, _, H, W = x.shape
        Ho, Wo = y.shape[2], y.shape[3]
        if need_gw:
            self._conv_wgrad(p, x, y, down, gy)
        if not need_gx:
            return None
        w_t = _flipped_transposed(w)                                                  # transpose of a stride-1 'same' conv
        if not down:
            return self.L.conv_same(gy, w_t)
        # stride 2 over the (0,1,0,1)-padded input: the data gradient is the stride-1 conv of the zero-interleaved output gradient,
        # placed at odd positions of an (H + 1) x (W + 1) canvas, cropped back to H x W
        canvas = torch.zeros((B, Co, H + 1, W + 1), device=gy.device, dtype=torch.float32)
        canvas[:, :, 1:2 * Ho:2, 1:2 * Wo:2] = gy
        return self.L.conv_same(canvas, w_t)[:, :, :H, :W]

    def _conv_wgrad(self, p, x, y, down, gy):
        w = self.sd[p + ".weight"]
        Co, Ci, k, _ = w.shape
        Ho, Wo = y.shape[2], y.shape[3]
        gy4 = _nhwc(gy)
        gyn = gy4.reshape(-1, Co)
        xn = _nhwc(x)
        G = self.L.wgrad_conv(xn, gy4, k, 2 if down else 1, 0 if down else k // 2) if hasattr(self.L, "wgrad_conv") else None
        if G is None:
            col = self.L.im2col(xn, k, 2 if down else 1, 0 if down else k // 2, Ho, Wo)
            G = self.L.gemm_tn(col, gyn)                                               # [k*k*Ci][Co], tap-major rows
        self._acc(p + ".weight", G.view(k * k, Ci, Co).permute(2, 1, 0).reshape(Co, Ci, k, k).contiguous())
        if (p + ".bias") in self.sd:
            self._acc(p + ".bias", self.L.colsum(gyn))

    # -- Normalize (+ swish) (encoder_decoder.py:29-35) -------------------------------------------------------------------------------
    def gn(self, p, x, swish):
        xn = _nhwc(x)
        y, stats = self.L.gn_fwd(xn, self.sd[p + ".weight"], self.sd[p + ".bias"], swish)
        self.ops.append(("gn", p, xn, stats, swish))
        return _nchw(y)

    def _gn_bwd(self, p, xn, stats, swish, gy):
        gx, dg, db = self.L.gn_bwd(xn, _nhwc(gy), stats, self.sd[p + ".weight"], self.sd[p + ".bias"], swish)
        self._acc(p + ".weight", dg)
        self._acc(p + ".bias", db)
        return _nchw(gx)

    # -- AttnBlock core (encoder_decoder.py:176-187) -----------------------------------------------------------------------------------
    def attention(self, q, k, v):
        o = self.L.attention(q, k, v)
        self.ops.append(("attn", q, k, v))
        return o

    def _attn_bwd(self, q, k, v, go):
        B, C, h, w = q.shape
        scale = float(int(C) ** (-0.5))
        gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        hw = (h, w)
        for b in range(B):
            Q, K, V, dO = (_nhwc(t[b:b + 1]).reshape(h * w, C) for t in (q, k, v, go))
            P = self.L.softmax_rows(self.L.gemm_nt(Q, K, hw).contiguous(), scale)       # w_ = softmax(scale q^T k)          :181-183
            dP = self.L.gemm_nt(dO, V, hw).contiguous()                                 # h_[n] = sum_j w_[n][j] v[j]        :186-187
            dS = self.L.softmax_bwd(P, dP, scale)
            dV = self.L.gemm_nt_ta(P, dO.t().contiguous(), hw)                          # P^T dO
            dQ = self.L.gemm_nt(dS, K.t().contiguous(), hw)
            dK = self.L.gemm_nt_ta(dS, Q.t().contiguous(), hw)                          # dS^T Q
            for dst, src in ((gq, dQ), (gk, dK), (gv, dV)):
                dst[b] = src.reshape(h, w, C).permute(2, 0, 1)
        return gq, gk, gv


class BlockGraph:
    """forward graph of the taming blocks over a Tape: a list of nodes (kind, output id, input ids, payload), ids index ``self.vals``; shared by
    the condition encoder (stage 2) and the deformable decoder (stage 3, decoder_train.py)"""

    def __init__(self, leaves, sd):
        self.L, self.sd = leaves, sd

    def _begin(self, *inputs):
        self.tape = Tape(self.L, self.sd)
        self.nodes, self.vals = [], list(inputs)

    def _push(self, kind, val, inputs, payload):
        self.vals.append(val)
        self.nodes.append((kind, len(self.vals) - 1, inputs, payload))
        return len(self.vals) - 1

    def _conv(self, p, i, down=False, need_gx=True, need_gw=True):
        y = self.tape.conv(p, self.vals[i], down, need_gx, need_gw)
        return self._push("op", y, (i,), self.tape.ops.pop())

    def _gn(self, p, i, swish):
        y = self.tape.gn(p, self.vals[i], swish)
        return self._push("op", y, (i,), self.tape.ops.pop())

    def _add(self, i, j):
        return self._push("add", self.vals[i] + self.vals[j], (i, j), None)

    def _fn(self, val, inputs, bwd):
        """a node whose backward rule is the closure ``bwd(gy) -> one gradient (or None) per input``"""
        return self._push("fn", val, tuple(inputs), bwd)

    def _resnet(self, p, i, shortcut="nin_shortcut"):
        """ResnetBlock.forward   encoder_decoder.py:117-137 (deformableDecoder_arch.py ResBlock :170-183 names its shortcut conv_out)"""
        h = self._conv(p + ".conv1", self._gn(p + ".norm1", i, True))
        h = self._conv(p + ".conv2", self._gn(p + ".norm2", h, True))
        s = self._conv(p + "." + shortcut, i) if (p + "." + shortcut + ".weight") in self.sd else i
        return self._add(s, h)

    def _attn(self, p, i):
        """AttnBlock.forward   encoder_decoder.py:168-192"""
        hn = self._gn(p + ".norm", i, False)
        q, k, v = (self._conv(p + "." + n, hn) for n in "qkv")
        o = self.tape.attention(self.vals[q], self.vals[k], self.vals[v])
        o = self._push("op", o, (q, k, v), self.tape.ops.pop())
        return self._add(i, self._conv(p + ".proj_out", o))

    def release(self):
        """drop the tape's activations now: the closures of `_fn` nodes refer back to the graph, so without this the memory of a finished step
        waits for the cyclic garbage collector (measured on stage 3: 33 GB peak and a caching allocator that keeps growing for a dozen steps)"""
        self.nodes, self.vals = [], []
        if getattr(self, "tape", None) is not None:
            self.tape.ops = []

    def _backprop(self, g):
        """``g``: {value id: gradient} seeds; walks the nodes in reverse; returns the gradients that reached the graph inputs"""
        T = self.tape

        def give(i, val):
            if val is not None:
                g[i] = g[i] + val if i in g else val

        while self.nodes:
            kind, out, inputs, op = self.nodes.pop()                    # popped: a node's saved activations are freed as soon as it is done
            gy = g.pop(out, None)
            self.vals[out] = None
            if gy is None:
                continue
            if kind == "add":
                give(inputs[0], gy)
                give(inputs[1], gy)
            elif kind == "fn":
                for i, val in zip(inputs, op(gy)):
                    give(i, val)
            elif op[0] == "conv":
                give(inputs[0], T._conv_bwd(op[1], op[2], op[3], op[4], op[5], gy, op[6]))
            elif op[0] == "gn":
                give(inputs[0], T._gn_bwd(op[1], op[2], op[3], op[4], gy))
            else:
                for i, val in zip(inputs, T._attn_bwd(op[1], op[2], op[3], gy)):
                    give(i, val)
        return g


class EncoderTrainer(BlockGraph):
    """forward with tape and backward of ConEncoder1; parameter names are the reference's state-dict keys under ``prefix`` ('RRDB')"""

    def __init__(self, leaves, sd, prefix="RRDB"):
        super().__init__(leaves, sd)
        self.p = prefix

    def forward(self, x):
        self._begin(x)
        sd, e = self.sd, self.p + ".encoder"
        h = self._conv(e + ".conv_in", 0, need_gx=False)
        for lvl in range(3):
            for blk in range(2):
                h = self._resnet("%s.down.%d.block.%d" % (e, lvl, blk), h)
                if ("%s.down.%d.attn.%d.q.weight" % (e, lvl, blk)) in sd:
                    h = self._attn("%s.down.%d.attn.%d" % (e, lvl, blk), h)
            if lvl != 2:
                h = self._conv("%s.down.%d.downsample.conv" % (e, lvl), h, down=True)
        h = self._resnet(e + ".mid.block_1", h)
        h = self._attn(e + ".mid.attn_1", h)
        h = self._resnet(e + ".mid.block_2", h)
        enc = self._conv(e + ".conv_out", self._gn(e + ".norm_out", h, True))
        cond_pre = self._conv(self.p + ".cond_conv.0", enc)
        color = self._conv(self.p + ".color_conv", enc)
        self.out_ids = (cond_pre, color)
        cond = torch.sigmoid(self.vals[cond_pre])                                       # ConditionEncoder.py:52
        return {"cond_feat": cond, "color_map": self.vals[color]}

    def backward(self, g_cond_feat, g_color_map):
        """gradients of the two heads -> {state-dict key: gradient} of every encoder parameter"""
        cond_pre, color = self.out_ids
        s = torch.sigmoid(self.vals[cond_pre])
        g = {cond_pre: g_cond_feat * s * (1.0 - s)}
        if g_color_map is not None:                  # None: the objective did not use color_map (mean = gt branch) -> color_conv gets no gradient
            g[color] = g_color_map
        self._backprop(g)
        return self.tape.grads


def stage2_step(sd, plan, lr, gt_latent, leaves, conv2d, flow_kernels=None, use_gt_mean=False):
    """One stage-2 objective evaluation with all gradients: -> (nll [B], {state-dict key: dL/dparam}) for L = nll.mean().
    ``lr`` preprocessed low-light input [B,3,H,W], ``gt_latent`` [B,3,H/4,W/4] (the frozen VQGAN's encoding of the ground truth).
    use_gt_mean: the Gaussian's mean is ``gt`` instead of the encoder's color_map -- the branch LLFlowVQGAN2.normal_flow takes when
    ``random.random() > opt['train_gt_ratio']`` is False (LLFlowVQGAN2_arch.py:109; 0.2 in train_stage2_LOL.yml); color_conv then receives no
    gradient (its keys are absent from the result, like the reference's ``param.grad is None``)."""
    enc = EncoderTrainer(leaves, sd)
    heads = enc.forward(lr)
    mean = gt_latent if use_gt_mean else heads["color_map"]
    nll, _, _, g_ft, g_mean, grads = flow_train.nll_forward_backward(plan, sd, gt_latent, heads["cond_feat"], mean, conv2d, kernels=flow_kernels)
    grads.update(enc.backward(g_ft, None if use_gt_mean else g_mean))
    return nll, grads


def draw_use_gt_mean(train_gt_ratio):
    """the reference's per-step draw (LLFlowVQGAN2_arch.py:109), consuming Python's global RNG exactly like it"""
    import random
    return not (random.random() > train_gt_ratio)


def _graphed_step(cfg, run, gt_latent, lr, params, use_gt_mean):
    """The whole objective + gradient evaluation (~2 500 kernel launches and torch ops at batch 4 x 320x320, CPU-launch bound when issued one by
    one) replayed as ONE CUDA graph per (shapes, branch, parameter storage).  The parameters are read in place at replay time -- the
    optimizer updates them in place, so their addresses are stable -- and everything derived from them (packed weights, flipped filters,
    the FlowPlan) is recomputed INSIDE the graph: the host-side caches are bypassed during capture.  The returned objective / gradients are the
    graph's static buffers, valid until the next call (``forward`` clones the objective; ``backward`` consumes the gradients before that)."""
    import contextlib
    from .engine import _Graphed
    key = (tuple(gt_latent.shape), tuple(lr.shape), bool(use_gt_mean), tuple(p.data_ptr() for p in params))
    graphs = cfg.setdefault("_graphs", {})
    g = graphs.get(key)
    if g is None:
        dense = getattr(cfg["leaves"], "dense", None)

        @contextlib.contextmanager
        def no_caches():
            global _NO_CACHE
            old = (_NO_CACHE, getattr(dense, "force_repack", False))
            _NO_CACHE = True
            if dense is not None:
                dense.force_repack = True
            try:
                yield
            finally:
                _NO_CACHE = old[0]
                if dense is not None:
                    dense.force_repack = old[1]

        if len(graphs) >= 4:
            graphs.clear()
        g = graphs[key] = _Graphed(None, run, gt_latent, lr, capture_ctx=no_caches)
    return g(gt_latent, lr)


class Stage2NLL(torch.autograd.Function):
    """The stage-2 objective as ONE autograd node over (gt_latent, lr, *parameters): ``forward`` evaluates the objective and every gradient
    with the library's kernels (stage2_step), ``backward`` hands the gradients to the parameters, rescaled by the incoming gradient (which
    must weigh the samples equally: ``nll.mean()``, ``nll.sum()``, a GradScaler factor ...).  This is what replaces
    ``z, nll, _ = netG(gt=..., lr=..., reverse=False); nll.mean().backward()`` of LLFlow_model.optimize_parameters (LLFlow_model.py:215-232)
    for a module whose parameters carry the reference's names."""

    @staticmethod
    def forward(ctx, cfg, gt_latent, lr, *params):
        keys = cfg["keys"]
        use_gt_mean = cfg.get("use_gt_mean", False)

        def run(gt, x):
            from . import flow
            sd = {k: p.detach() for k, p in zip(keys, params)}
            sd.update(cfg.get("buffers", {}))
            plan = flow.FlowPlan(sd, gt.device)
            nll, grads = stage2_step(sd, plan, x, gt, cfg["leaves"], cfg["conv2d"], flow_kernels=cfg.get("flow_kernels"), use_gt_mean=use_gt_mean)
            return nll, [grads.get(k) for k in keys]

        with torch.no_grad():
            if cfg.get("graph") and gt_latent.is_cuda:
                nll, glist = _graphed_step(cfg, run, gt_latent, lr, params, use_gt_mean)
            else:
                nll, glist = run(gt_latent, lr)
        ctx.batch = gt_latent.shape[0]
        ctx.has = [g is not None for g in glist]
        ctx.save_for_backward(*[g for g in glist if g is not None])
        return nll.clone() if cfg.get("graph") else nll

    @staticmethod
    def backward(ctx, g_nll):
        if not bool((g_nll == g_nll[0]).all()):
            raise RuntimeError("Stage2NLL supports reductions that weigh every sample equally (nll.mean(), nll.sum())")
        scale = g_nll[0] * ctx.batch                                     # stored gradients are those of nll.mean()
        try:
            scaled = torch._foreach_mul(list(ctx.saved_tensors), scale)  # one fused launch instead of one per parameter (628 of them)
        except (RuntimeError, TypeError):
            scaled = [g * scale for g in ctx.saved_tensors]
        saved = iter(scaled)
        return (None, None, None) + tuple(next(saved) if h else None for h in ctx.has)


def stage2_nll(named_parameters, gt_latent, lr, leaves, conv2d, flow_kernels=None, train_gt_ratio=0.0, use_gt_mean=None, graph=False):
    """``named_parameters``: iterable of (state-dict key, Parameter) of the generator (netG of train_stage2.py: ``RRDB.*`` and
    ``flowUpsamplerNet.*``).  Returns the per-sample objective [B] with the graph edge to every parameter, so that
    ``stage2_nll(...).mean().backward()`` fills ``param.grad`` exactly like the reference's autograd.  ``train_gt_ratio``
    (opt['train_gt_ratio']): probability of the ``mean = gt`` branch, drawn per call like the reference; ``use_gt_mean`` overrides the draw.
    ``graph=True``: the evaluation replays as one CUDA graph (see _graphed_step); pass the SAME ``cfg`` holder via ``graph`` (a dict) to keep
    the captured graphs across calls."""
    named = [(k, p) for k, p in named_parameters]
    if use_gt_mean is None:
        use_gt_mean = draw_use_gt_mean(train_gt_ratio)
    cfg = graph if isinstance(graph, dict) else {}
    cfg.update({"keys": [k for k, _ in named], "leaves": leaves, "conv2d": conv2d, "flow_kernels": flow_kernels, "use_gt_mean": use_gt_mean,
                "graph": bool(graph) or isinstance(graph, dict)})
    return Stage2NLL.apply(cfg, gt_latent, lr, *[p for _, p in named])
