"""Dense-operator backend of the engine (tcgen05 kernels of libglare_b200.so) and its roofline bookkeeping (see engine.py docstring).

``make_dense("auto")``        = "tc-bf16x3": fp32-grade (each fp32 operand split into two bf16 pieces, three bf16 passes; measured as
                              accurate as 3xTF32 on this network, profiles/r01n_fullsize_parity_420x620.txt)
``make_dense("tc-tf32bf16x2" | "tc-3xtf32" | "tc-tf32" | "tc-bf16")``  the same kernels in the other operand modes

There is no library (cuDNN / cuBLAS) backend and no fallback: a shape outside the kernels' coverage raises.
"""
import os
import weakref

import torch
import torch.nn.functional as F


def dcn_flops(C, Cout, px):
    return 2.0 * C * Cout * 9 * px


class _Bracket:
    """CUDA-event bracket (current stream) recorded into a timers dict: name -> [(e0, e1, algorithmic_flops)]"""

    def __init__(self, timers, name, flops=0.0):
        self.timers, self.name, self.flops = timers, name, flops

    def __enter__(self):
        if self.timers is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.timers is not None:
            self.e1.record()
            self.timers.setdefault(self.name, []).append((self.e0, self.e1, self.flops))


class Operand:
    """Tensor-core operand of a conv: NHWC ``hi`` (+ ``lo`` for the 3xTF32 split) of a logical [B,C,H,W] activation."""

    def __init__(self, mode, hi, lo, B, C, H, W):
        self.mode, self.hi, self.lo, self.B, self.C, self.H, self.W = mode, hi, lo, B, C, H, W

    def dense(self):
        """fp32 logical-NCHW (channels_last) value of the operand: hi + lo is exact for the tf32 split"""
        if self.mode == 4:                       # interleaved bf16 pair: x = a1 + a2
            pair = self.hi.view(-1, 2, 32).float()
            v = (pair[:, 0] + pair[:, 1]).reshape(self.B, self.H, self.W, self.C)
        elif self.lo is None:
            v = self.hi.float()
        elif self.lo.dtype == torch.float32:
            v = self.hi + self.lo
        else:                                    # mode 3: interleaved bf16 x tensor, second half of every 64 = bf16(x - hi)
            v = self.hi + self.lo.view(-1, 2, 32)[:, 1].float().reshape(self.hi.shape)
        return v.view(self.B, self.H, self.W, self.C).permute(0, 3, 1, 2)


def _nhwc(x):
    """logical [B,C,H,W] fp32 tensor -> the same storage viewed as contiguous NHWC (converting layout if needed)"""
    if x.dtype != torch.float32:
        x = x.float()
    if not x.is_contiguous(memory_format=torch.channels_last) or (x.shape[1] > 1 and x.stride(1) != 1):
        x = x.contiguous(memory_format=torch.channels_last)
    return x.permute(0, 2, 3, 1)


class TcDense:
    """tcgen05 implicit-GEMM convolutions, attention GEMMs, DCNv2 and the GroupNorm/swish operand kernels (libglare_b200.so).

    mode 0 bf16 operands, 1 tf32, 2 3xTF32, 3 tf32 + 2 x bf16 cross terms, 4 bf16x3 (default; fp32-grade).  Every conv shape of the path
    runs on the tcgen05 kernel (3-channel layers through channel padding, Downsample through the TMA traversal stride, Upsample as four
    sub-pixel phases) and so do the attention matmuls.  A shape outside the kernels' coverage raises NotImplementedError (no library
    fallback); ``self.fallbacks`` only records the fused-softmax -> exact-softmax switch of the attention path."""

    def __init__(self, mode):
        from . import ops
        self.ops = ops
        self.mode = mode
        self.name = {0: "tcgen05-bf16", 1: "tcgen05-tf32", 2: "tcgen05-3xtf32", 3: "tcgen05-tf32+2xbf16", 4: "tcgen05-bf16x3"}[mode]
        self.dtype_name = {0: "bf16", 1: "tf32", 2: "fp32 (3xTF32 tensor-core emulation, fp32 accumulate)",
                           3: "fp32 (tf32 x tf32 + two bf16 cross terms on tensor cores, fp32 accumulate)",
                           4: "fp32 split into two bf16 pieces per operand, three bf16 tensor-core passes (16-bit significand), fp32 accumulate"}[mode]
        self.bke = 64 if mode == 0 else 32
        self._w = {}
        self.force_repack = False      # CUDA-graph capture of a TRAINING step: every replay must repack from the current weights
        self.fallbacks = {}
        # GroupNorm sum / sum of squares of a conv's output taken in its epilogue (column sums of the staged output tile -> per-(tile, warp)
        # partials -> fp64 finish launch; no atomics, no shuffles over pixels), which removes the statistics pass over the activation.
        # Round 1's version (warp reductions + fp64 atomics on the TMEM-drain path) cost more than it saved and was off.
        self.fuse_gn_stats = not os.environ.get("GLARE_NO_FUSE_GN_STATS")      # A/B switch
        self.dcn_tc = True             # DCNv2 on the tensor-core kernel (dcn_tc.cu); False -> fp32 FMA kernel (dcn.cu)
        self.attn_s_budget = None      # bytes of score / operand matrix materialised per pass; None: sized from free HBM at first use (_budget)
        # mode 4: softmax fused into the epilogues of the two GEMMs (csrc/attn.cu): exp against a Cauchy-Schwarz row reference in the scores
        # GEMM, 1 / row sum in the P V GEMM; rows that fall outside the safe window raise `attn_flag` and `attention_verified()` tells
        # the caller (engine.infer) to re-run with the exact three-kernel path
        self.attn_fused = mode in (0, 4) and not os.environ.get("GLARE_ATTN_UNFUSED")      # (mode 0: P~ written as the single-piece bf16 operand)
        self.attn_margin = 60.0
        # row reference of the fused softmax: "sampled" = maximum over 128 strided keys + 50 (one small extra GEMM per block; robust to loose
        # |q||k| bounds), "cauchy" = the Cauchy-Schwarz bound - margin of round 1 (needs no extra GEMM; trips when the bound is > ~115 above
        # the row maximum)
        self.attn_ref = os.environ.get("GLARE_ATTN_REF", "sampled")
        self.attn_ref_keys, self.attn_ref_offset = 128, 50.0
        # keys contracted per P V launch: longer rows (1080p: 131 648 keys) run in bands chained through the residual epilogue, which keeps the
        # tensor-core accumulator's truncation bias at the level of the 600x400 shape (16 320 keys: one launch)
        self.attn_key_band = 16384
        self.attn_flag = None
        self.pack_epilogue = mode == 4 and not os.environ.get("GLARE_NO_PACK_EPILOGUE")   # A/B switch: operands written by the producing conv
        self.timers = None             # bench.py: dict name -> [(start_event, end_event, algorithmic_flops)]
        self.last_timers, self.last_steps = None, 1

    def _t(self, name, flops=0.0):
        return _Bracket(self.timers, name, flops)

    def _cached(self, key, w, make):
        """packed-weight cache: one entry per (storage address, shape, role), REPLACED when the tensor's in-place version moves (an optimizer
        step bumps it on every weight: keying on the version would add a packed copy of every weight per training step).  Entries hold the
        source tensor by weak reference: a different tensor object at a recycled address is repacked, and entries whose source has died
        (per-step temporaries of the training path: flipped filters, the flow's hoisted weight) are swept when new keys arrive."""
        ent = self._w.get(key)
        if self.force_repack or ent is None or ent[0]() is not w or ent[1] != w._version:
            if ent is None and len(self._w) >= 256:
                for k in [k for k, e in self._w.items() if e[0]() is None]:
                    del self._w[k]
            ent = self._w[key] = (weakref.ref(w), w._version, make())
        return ent[2]

    def _weights(self, w, pad_c=0):
        def make():
            wp = F.pad(w, (0, 0, 0, 0, 0, pad_c)) if pad_c else w            # zero input channels up to the K chunk
            return self.ops.conv_pack_weight(self.mode, wp)
        return self._cached((w.data_ptr(), tuple(w.shape), pad_c), w, make)

    def _operand(self, x):
        """logical [B,C,H,W] fp32 tensor -> Operand, input channels zero-padded to a multiple of the 128-byte K chunk"""
        if isinstance(x, Operand):
            return x, 0
        with self._t("prep_act"):
            xn = _nhwc(x)
            pad_c = (-xn.shape[3]) % self.bke
            if pad_c:
                xn = F.pad(xn, (0, pad_c))
            B, H, W, C = xn.shape
            hi, lo = self.ops.conv_prep_act(self.mode, xn)
        return Operand(self.mode, hi, lo, B, C, H, W), pad_c

    @staticmethod
    def _unsupported(what):
        raise NotImplementedError("glare_b200: %s is outside the tcgen05 kernels' coverage and there is no library fallback" % what)

    def conv2d(self, x, w, b=None, stride=1, padding=1, residual=None, gn_stats=True):
        """gn_stats=False: the output does not feed a Normalize (training backward convs): skip the epilogue statistics"""
        Cin = x.C if isinstance(x, Operand) else x.shape[1]
        Cout, ks = w.shape[0], w.shape[2]
        shape_ok = w.shape[2] == w.shape[3] and ks in (1, 3) and stride == 1 and padding == ks // 2
        if not shape_ok:
            self._unsupported("conv %dx%d %d->%d stride %d padding %d" % (ks, w.shape[3], Cin, Cout, stride, padding))
        op, pad_c = self._operand(x)
        w_hi, w_lo = self._weights(w, pad_c)
        flops = 2.0 * op.B * op.H * op.W * Cin * Cout * ks * ks
        if Cout % 4:                                   # 3-channel heads: padded output pixel stride, sliced view returned
            ldy = (Cout + 3) // 4 * 4
            y = torch.empty((op.B, op.H, op.W, ldy), device=op.hi.device, dtype=torch.float32)
            with self._t("conv_tc", flops):
                self.ops.conv2d_nhwc_tc_ex(self.mode, op.hi, op.lo, w_hi, w_lo, y, op.B, op.H, op.W, op.C, Cout, ldy, 0, bias=b, ksize=ks)
            y = y[..., :Cout].permute(0, 3, 1, 2)
            return y if residual is None else y + residual
        res = _nhwc(residual) if residual is not None else None
        y = torch.empty((op.B, op.H, op.W, Cout), device=op.hi.device, dtype=torch.float32)
        stats = self._new_stats(op.B, Cout) if gn_stats else None
        with self._t("conv_tc", flops):
            self.ops.conv2d_nhwc_tc_g(self.mode, 0, op.hi, op.lo, w_hi, w_lo, b, res, y, op.B, op.H, op.W, op.C, Cout, ksize=ks,
                                      gn_stats=stats)
        return self._tag(y.permute(0, 3, 1, 2), stats)

    def conv2d_operand(self, x, w, b=None, row_sq=False):
        """stride-1 'same' conv whose only consumer is another tensor-core GEMM: the epilogue writes the bf16x3 operand directly (no fp32
        tensor, no conversion pass).  row_sq: also per-row partial sums of squares (AttnBlock q / k -> the fused softmax's row reference).
        None when this backend / shape has no such epilogue (the caller falls back to conv2d)."""
        Cout, ks = w.shape[0], w.shape[2]
        Cin = x.C if isinstance(x, Operand) else x.shape[1]
        if self.mode != 4 or not self.pack_epilogue or Cout % 32 or Cin % self.bke or w.shape[2] != w.shape[3] or ks not in (1, 3):
            return None
        op, pad_c = self._operand(x)
        w_hi, _ = self._weights(w, pad_c)
        with self._t("conv_tc", 2.0 * op.B * op.H * op.W * Cin * Cout * ks * ks):
            y, sq = self.ops.conv2d_nhwc_tc_pack(self.mode, op.hi, w_hi, b, op.B, op.H, op.W, op.C, Cout, ks, row_sq=row_sq)
        out = Operand(self.mode, y, None, op.B, Cout, op.H, op.W)
        out.row_sq = sq
        return out

    def cat_operand(self, a, b):
        """operand of torch.cat([a, b], dim=1) for two fp32 activations (WarpBlock's offset-conv input): one pass, no fp32 cat tensor.
        None when this mode / shape has no such kernel (the caller concatenates)."""
        if self.mode != 4 or a.shape[1] % 32 or b.shape[1] % 32 or a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
            return None
        with self._t("prep_act"):
            an, bn = _nhwc(a), _nhwc(b)
            B, H, W, Ca = an.shape
            hi = self.ops.aft_cat_operand(an.contiguous(), bn.contiguous())
        return Operand(self.mode, hi, None, B, Ca + bn.shape[3], H, W)

    def _new_stats(self, B, Cout):
        """fp64 [B,32,2] buffer for GroupNorm statistics fused into the conv epilogue (None when the output cannot feed Normalize)"""
        if not self.fuse_gn_stats or Cout % 128 or Cout > 512:
            return None
        return torch.empty((B, 32, 2), device="cuda", dtype=torch.float64)

    @staticmethod
    def _tag(y, stats):
        if stats is not None:
            y._glare_gn_stats = stats          # consumed by gn_swish if this very tensor is normalised next
        return y

    def upsample_conv(self, x, w, b=None):
        """Upsample.forward (encoder_decoder.py:49-53): nearest x2 + 3x3 conv, evaluated on the LOW-resolution input as four
        sub-pixel phases with pre-summed 2x2 filters (no upsampled copy, 4/9 of the FLOPs)"""
        if tuple(w.shape[2:]) != (3, 3) or w.shape[0] % 4 or x.shape[1] % self.bke:
            self._unsupported("Upsample conv %s" % (tuple(w.shape),))
        def make():
            rows = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}          # phase -> which 3x3 taps land on source row/col 0 and 1
            ph = {}
            for a in (0, 1):
                for c in (0, 1):
                    w2 = torch.stack([torch.stack([w[:, :, list(rows[a][r])][:, :, :, list(rows[c][q])].sum(dim=(2, 3))
                                                   for q in (0, 1)], dim=-1) for r in (0, 1)], dim=-2)      # [Co,Ci,2,2]
                    ph[(a, c)] = self.ops.conv_pack_weight(self.mode, w2.contiguous())
            return ph
        phases = self._cached(("up2", w.data_ptr(), tuple(w.shape)), w, make)
        op, _ = self._operand(x)
        Cout = w.shape[0]
        y = torch.empty((op.B, 2 * op.H, 2 * op.W, Cout), device=op.hi.device, dtype=torch.float32)
        with self._t("conv_tc", 2.0 * op.B * op.H * op.W * op.C * Cout * 16):
            for (pa, pb), (w_hi, w_lo) in phases.items():
                self.ops.conv2d_nhwc_tc_g(self.mode, 2, op.hi, op.lo, w_hi, w_lo, b, None, y, op.B, op.H, op.W, op.C, Cout, ksize=2, pa=pa,
                                          pb=pb)
        return y.permute(0, 3, 1, 2)            # (the sub-pixel phases store directly: no staged tile, statistics by the separate kernel)

    def downsample_conv(self, x, w, b=None, gn_stats=True):
        """Downsample.forward (encoder_decoder.py:68-72): pad (0,1,0,1) + 3x3 stride-2 conv, padding by TMA zero fill"""
        if tuple(w.shape[2:]) != (3, 3) or w.shape[0] % 4:
            self._unsupported("Downsample conv %s" % (tuple(w.shape),))
        op, pad_c = self._operand(x)
        w_hi, w_lo = self._weights(w, pad_c)
        Ho, Wo = (op.H - 2) // 2 + 1, (op.W - 2) // 2 + 1
        Cout = w.shape[0]
        y = torch.empty((op.B, Ho, Wo, Cout), device=op.hi.device, dtype=torch.float32)
        stats = self._new_stats(op.B, Cout) if gn_stats else None
        with self._t("conv_tc", 2.0 * op.B * Ho * Wo * x.shape[1] * Cout * 9):
            self.ops.conv2d_nhwc_tc_g(self.mode, 1, op.hi, op.lo, w_hi, w_lo, b, None, y, op.B, op.H, op.W, op.C, Cout, gn_stats=stats)
        return self._tag(y.permute(0, 3, 1, 2), stats)

    def gn_swish(self, x, gamma, beta, swish=True):
        C = x.shape[1]
        if C % 128 != 0:
            self._unsupported("GroupNorm(32) over %d channels" % C)
        with self._t("groupnorm"):
            stats = getattr(x, "_glare_gn_stats", None)            # produced by the conv epilogue that wrote x
            xn = _nhwc(x)
            B, H, W, _ = xn.shape
            if stats is None:
                stats = self.ops.gn_stats(xn, B, H * W, C)
            hi, lo = self.ops.gn_apply(self.mode, xn, stats, gamma, beta, swish, B, H * W, C)
        return Operand(self.mode, hi, lo, B, C, H, W)

    def dcn_pack(self, x, offmask_raw, weight, bias, dg):
        """DCNv2Pack.forward after conv_offset (deformableDecoder_arch.py:141-152): x [B,C,H,W], raw conv_offset output
        [B,27*dg,H,W] -> y [B,Cout,H,W].  None when the shape is outside the tensor-core kernel's coverage."""
        C, Cout = x.shape[1], weight.shape[0]
        if tuple(weight.shape[2:]) != (3, 3) or C % dg or C % self.bke or (C // dg) % 8 or Cout % 4 or not self.dcn_tc:
            return None
        with self._t("dcn_tc", 2.0 * x.shape[0] * x.shape[2] * x.shape[3] * C * Cout * 9):
            xn, on = _nhwc(x), _nhwc(offmask_raw)
            B, H, W, _ = xn.shape
            w_hi, w_lo = self._weights(weight)
            y = self.ops.dcnv2_pack_fwd_nhwc_tc(self.mode, xn, on, w_hi, w_lo, bias, B, H, W, C, Cout, dg)
        return y.permute(0, 3, 1, 2)

    def _budget(self):
        """bytes of N x N operand per pass.  600x400 needs 1.06 GB per sample (one pass).  1080p (131 648 tokens, 69 GB per sample) runs in
        bands of query rows: the band must hold enough 256-row tiles to fill the 74 CTA pairs of the P V GEMM (N = 512 -> two output
        blocks), so the budget is a quarter of the free HBM up to 16 GiB (31 K query rows at 1080p; round 1's fixed 1.5 GiB gave 3.9 K rows
        = 30 work items for 74 pairs)."""
        if self.attn_s_budget is None:
            free, _ = torch.cuda.mem_get_info()
            self.attn_s_budget = int(max(3 << 29, min(16 << 30, free // 4)))
        return self.attn_s_budget

    def attention(self, q, k, v, as_operand=False, fused=None):
        """AttnBlock core (encoder_decoder.py:176-187) on the tcgen05 GEMM path: per sample and per band of query rows
        S = Q K^T (W = K[n]) -> fused scale + row softmax emitting the operand P -> O = P V (W = V[n]^T)."""
        q_is_op, k_is_op = isinstance(q, Operand), isinstance(k, Operand)
        C = q.C if q_is_op else q.shape[1]
        if C % self.bke != 0:
            self._unsupported("attention over %d channels" % C)
        ops = self.ops
        with self._t("attention_total(incl. its conv_tc GEMMs)"):
            vn = _nhwc(v)
            B, h, w, _ = vn.shape
            N = h * w
            Np = (N + 63) // 64 * 64
            qn = None if q_is_op else _nhwc(q)
            kn = None if k_is_op else _nhwc(k)
            q_hi, q_lo = (q.hi, q.lo) if q_is_op else ops.conv_prep_act(self.mode, qn)
            k_hi, k_lo = (k.hi, k.lo) if k_is_op else ops.conv_prep_act(self.mode, kn)      # K[n] as GEMM weights [N][C]
            vt_hi, vt_lo = ops.attn_transpose_v(self.mode, vn, B, N, C, Np)
            if (self.attn_fused if fused is None else (fused and self.mode in (0, 4))) and N >= 32:
                sq = (getattr(q, "row_sq", None), getattr(k, "row_sq", None))
                if as_operand and self.pack_epilogue and C % 32 == 0:  # the P V epilogue writes proj_out's operand
                    out_op = ops._hi_alloc(self.mode, (B, h, w, C), vn.device)
                    self._attention_fused(qn, kn, q_hi, k_hi, vt_hi, out_op, B, h, w, N, Np, C, sq, pack=True)
                    return Operand(self.mode, out_op, None, B, C, h, w)
                out = torch.empty((B, h, w, C), device=vn.device, dtype=torch.float32)
                self._attention_fused(qn, kn, q_hi, k_hi, vt_hi, out, B, h, w, N, Np, C, sq)
                return out.permute(0, 3, 1, 2)
            out = torch.empty((B, h, w, C), device=vn.device, dtype=torch.float32)
            # query rows per pass: the whole sample when its score matrix fits the budget (tile-filling GEMMs matter more
            # than L2 residency of S: the P V GEMM needs >= 74 M-tiles to give every SM a 128x256 tile), else 8-row multiples
            budget = self._budget() // 2                      # S (fp32) and P (operand) are both materialised on this path
            band = h if N * Np * 4 <= budget else max(8, (budget // (w * Np * 4)) // 8 * 8)
            rows_max = min(band, h) * w
            S = torch.empty((rows_max, Np), device=vn.device, dtype=torch.float32)
            p_hi = ops._hi_alloc(self.mode, (rows_max, Np), vn.device)
            p_lo = ops._lo_like(self.mode, p_hi)
            scale = float(int(C) ** (-0.5))
            for b in range(B):
                for r0 in range(0, h, band):
                    r1 = min(h, r0 + band)
                    bh = r1 - r0
                    gemm_flops = 2.0 * bh * w * N * C
                    with self._t("conv_tc", gemm_flops):          # the same tcgen05 kernel: counted with the convolutions
                        ops.conv2d_nhwc_tc_ex(self.mode, q_hi[b, r0:r1], None if q_lo is None else q_lo[b, r0:r1], k_hi[b],
                                              None if k_lo is None else k_lo[b], S, 1, bh, w, C, N, Np, 0)
                    with self._t("attn_softmax"):
                        ops.attn_softmax_rows(self.mode, S, bh * w, Np, N, Np, scale, p_hi, p_lo, Np)
                    with self._t("conv_tc", gemm_flops):
                        if self.mode in (0, 4):           # key bands for very long rows, as on the fused path (no row scale: P is normalised)
                            ops.attn_pv_tc(self.mode, p_hi, vt_hi[b], None, out[b, r0:r1], bh, w, Np, C, C, key_band=self.attn_key_band)
                        else:
                            ops.conv2d_nhwc_tc_ex(self.mode, p_hi, p_lo, vt_hi[b], None if vt_lo is None else vt_lo[b], out[b, r0:r1],
                                                  1, bh, w, Np, C, C, 0)
        return out.permute(0, 3, 1, 2)

    def _attention_fused(self, qn, kn, q_op, k_op, vt_op, out, B, h, w, N, Np, C, sq=(None, None), pack=False):
        """mode 4: scores GEMM with the exp epilogue -> row-sum finish -> P V GEMM with the 1 / row-sum epilogue (csrc/attn.cu).
        qn / kn: fp32 NHWC q / k (row norms by a pass over them) or None when sq carries the partial sums of squares their conv left (only the
        "cauchy" reference needs either)."""
        ops, dev = self.ops, q_op.device
        if self.attn_flag is None or self.attn_flag.device != dev:
            self.attn_flag = torch.zeros((1,), device=dev, dtype=torch.int32)
        scale = float(int(C) ** (-0.5))
        q_norm = torch.empty((B, N), device=dev, dtype=torch.float32)
        if self.attn_ref == "sampled":
            # scores of every query against n_sub strided keys of its sample (per-sample weights: one launch for the batch) -> row maximum
            k_max = None
            n_sub = min(N, self.attn_ref_keys)
            sel = torch.linspace(0, N - 1, n_sub, device=dev).round().long()
            k_sub = k_op.view(B, N, -1).index_select(1, sel).contiguous()            # [B][n_sub][2C] operand rows = GEMM weights
            ld = (n_sub + 3) // 4 * 4
            s_sub = torch.empty((B * N, ld), device=dev, dtype=torch.float32)
            with self._t("conv_tc", 2.0 * B * N * n_sub * C):
                ops.conv2d_nhwc_tc_ex(self.mode, q_op, None, k_sub, None, s_sub, B, h, w, C, n_sub, ld, n_sub * C)
            with self._t("attn_softmax"):
                ops.attn_row_ref(s_sub, B * N, ld, n_sub, scale, self.attn_ref_offset, q_norm)
        else:
            k_max = torch.zeros((B,), device=dev, dtype=torch.int32)
            with self._t("attn_softmax"):
                if sq[0] is not None:
                    ops.attn_row_norm_finish(sq[0][0], sq[0][1], B * N, N, norm_out=q_norm)
                else:
                    ops.attn_row_norm(qn, B * N, C, N, norm_out=q_norm)
                if sq[1] is not None:
                    ops.attn_row_norm_finish(sq[1][0], sq[1][1], B * N, N, max_bits=k_max)
                else:
                    ops.attn_row_norm(kn, B * N, C, N, max_bits=k_max)
        # whole-sample passes while the operand matrix fits the budget, else bands of 8-row multiples
        budget = self._budget()
        band = h if N * Np * 4 <= budget else max(8, (budget // (w * Np * 4)) // 8 * 8)
        rows_max = min(band, h) * w
        p_op = ops._hi_alloc(self.mode, (rows_max, Np), dev)
        n32 = (N + 31) // 32 * 32
        if self.mode == 4 and n32 < Np:
            p_op[:, 2 * n32:].zero_()                                    # operand chunks past the last key chunk are never written
        # (mode 0: the epilogue writes whole 64-key rows, zeros for keys >= N, and Np is the next multiple of 64)
        part = torch.empty(((N + 63) // 64, rows_max), device=dev, dtype=torch.float32)
        row_scale = torch.empty((rows_max,), device=dev, dtype=torch.float32)
        for b in range(B):
            for r0 in range(0, h, band):
                r1 = min(h, r0 + band)
                bh = r1 - r0
                gemm_flops = 2.0 * bh * w * N * C
                with self._t("conv_tc", gemm_flops):
                    nb = ops.attn_scores_exp_tc(self.mode, q_op[b, r0:r1], k_op[b], bh, w, C, N, Np, scale, self.attn_margin,
                                                q_norm[b, r0 * w:], None if k_max is None else k_max[b:], p_op, part, rows_max)
                with self._t("attn_softmax"):
                    ops.attn_row_sum_finish(part, rows_max, nb, bh * w, row_scale, self.attn_flag)
                with self._t("conv_tc", gemm_flops):
                    ops.attn_pv_tc(self.mode, p_op, vt_op[b], row_scale, out[b, r0:r1], bh, w, Np, C, C, pack_out=pack,
                                   key_band=self.attn_key_band)

    def attention_verified(self):
        """True when every fused-softmax row so far stayed inside the safe window (one 4-byte device read).  On False the fused path is
        switched off for this backend and the caller must recompute with the exact path."""
        if self.attn_flag is None or not self.attn_fused:
            return True
        if int(self.attn_flag.item()) == 0:
            return True
        self.attn_fused = False
        self.attn_flag.zero_()
        self.fallbacks["attention: fused softmax window exceeded -> exact softmax path"] = 1
        return False

    def breakdown(self, eng_timers, steps):
        out = {}
        for src in (self.last_timers or {}, eng_timers or {}):
            for k, ev in src.items():
                if ev and ev[0][0] is not None:
                    out[k] = round(sum(e[0].elapsed_time(e[1]) for e in ev) / max(1, steps), 2)
        return out

    def roofline(self, timers, eng, B, lr_shape, pk):
        """Dominant kernel: conv_tc_kernel (all tensor-core conv launches of one step).  Algorithmic work =
        2*Cin*Cout*k*k FLOP per output pixel, summed over the launches; the 3xTF32 mode issues 3 MMAs
        per algorithmic MAC, the roofline counts the algorithmic ones."""
        ev = (self.last_timers or {}).get("conv_tc", [])
        if not ev:
            return None
        ms = sum(a.elapsed_time(b) for a, b, _ in ev)
        fl = sum(f for _, _, f in ev)
        steps = max(1, self.last_steps)
        ach = fl / (ms / 1e3) / 1e12
        issued = {0: 0.5, 1: 1.0, 2: 3.0, 3: 2.0, 4: 1.5}[self.mode]  # tf32-equivalent MMA passes per algorithmic MAC
        return {"kernel": "conv_tc_kernel (tcgen05 implicit GEMM: convolutions + attention GEMMs, %d launches/step, mode %s)"
                          % (len(ev) // steps, self.name),
                "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
                "traffic": self._traffic(B), "ms_per_step": ms / steps, "share_of_step": None,
                "algorithmic_tflop_per_step": fl / steps / 1e12, "launches_per_step": len(ev) // steps,
                "mma_passes_per_algorithmic_mac": issued,
                "issued_bf16_equivalent_tflops": ach * issued * 2.0, "issued_frac_of_peak": ach * issued * 2.0 / pk["tensor"],
                "note": "achieved counts algorithmic FLOPs (2*Cin*Cout*k*k per pixel, 4*N*N*C per attention block); this mode issues %g "
                        "tf32-equivalent (= %g bf16) tensor-core passes per algorithmic MAC, the peak is the bf16 one" % (issued, 2 * issued),
                "peak_source": pk["src"] + " (cuBLAS bf16 sustained)"}

    def _traffic(self, B):
        """dram__bytes_read.sum + dram__bytes_write.sum per launch of conv_tc_kernel from the committed ncu capture of this configuration
        (profiles/conv_tc_traffic.json, written by tools/traffic_summary.py on the GPU box).  Only returned when the capture was taken from
        the SAME kernel sources (sha256 of csrc/, glare_b200.build.source_hash) in the same mode and batch; otherwise None."""
        import json
        from .build import source_hash
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "conv_tc_traffic.json")
        try:
            with open(path) as f:
                t = json.load(f)
            if t.get("mode") != self.name or t.get("batch") != B or t.get("csrc_sha16") != source_hash():
                return None
            return t["dram_bytes_per_launch"]
        except Exception:
            return None


def make_dense(name="auto"):
    modes = {"auto": 4, "tc-bf16x3": 4, "tc-tf32bf16x2": 3, "tc-3xtf32": 2, "tc-tf32": 1, "tc-bf16": 0}
    if name not in modes:
        raise ValueError("unknown dense backend %r (tcgen05 operand modes: %s)" % (name, ", ".join(sorted(modes))))
    return TcDense(modes[name])
