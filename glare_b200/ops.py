"""Operator-level wrappers: allocate outputs with torch, call the C ABI on the current stream.

Signatures follow the reference operators they replace (cited per function; paths relative to
/root/reference/code/models/modules).
"""
import torch

from . import _lib
from ._lib import f32c, lib, ptr, require_cuda, stream

LAUNCHES = 0          # kernels of libglare_b200.so launched by this process (every C-ABI call below launches exactly one)


def check(code, what):
    global LAUNCHES
    _lib.check(code, what)
    LAUNCHES += 1


# ------------------------------------------------------------------------------------------- VQ
def vq_pack_codebook(codebook):
    """codebook [K,3] -> packed [K,4] {e0,e1,e2,|e|^2}  (|e|^2 rounded as quantize.py:281 does on CPU)."""
    require_cuda(codebook)
    cb = f32c(codebook)
    if cb.dim() != 2 or cb.shape[1] != 3:
        raise ValueError("VQ kernel supports e_dim == 3 (LOL.yml network_VQGAN embed_dim); got %s" % (tuple(cb.shape),))
    packed = torch.empty((cb.shape[0], 4), device=cb.device, dtype=torch.float32)
    check(lib().glare_vq_pack_codebook_f32(ptr(cb), cb.shape[0], ptr(packed), stream()), "glare_vq_pack_codebook_f32")
    return packed


def vq_lookup(z, packed_codebook):
    """VectorQuantizer2.forward core (quantize.py:276-301): z [B,3,h,w] -> (indices int64 [B*h*w], z_q [B,3,h,w])."""
    require_cuda(z, packed_codebook)
    z = f32c(z)
    if z.dim() != 4 or z.shape[1] != 3:
        raise ValueError("expected z [B,3,h,w], got %s" % (tuple(z.shape),))
    B, _, h, w = z.shape
    idx = torch.empty((B * h * w,), device=z.device, dtype=torch.int64)
    zq = torch.empty_like(z)
    check(lib().glare_vq_argmin_gather_f32(ptr(z), ptr(packed_codebook), B, h * w, packed_codebook.shape[0], ptr(idx),
                                           ptr(zq), stream()), "glare_vq_argmin_gather_f32")
    return idx, zq


# ------------------------------------------------------------------------------------------- DCN
def dcn_pack_weight(weight):
    """[Cout,C,kh,kw] -> [kh*kw][C][Cout] (once per weight update)."""
    require_cuda(weight)
    w = f32c(weight)
    Co, C, kh, kw = w.shape
    out = torch.empty((kh * kw, C, Co), device=w.device, dtype=torch.float32)
    check(lib().glare_dcn_pack_weight_f32(ptr(w), Co, C, kh, kw, ptr(out), stream()), "glare_dcn_pack_weight_f32")
    return out


def modulated_deform_conv(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                          deformable_groups=1, packed_weight=None):
    """ModulatedDeformConvFunction.forward (ops/dcn/deform_conv.py:124-153), same argument order and meaning.
    Inference only (no autograd graph is recorded)."""
    if not input.is_cuda:
        raise NotImplementedError            # deform_conv.py:143-144
    require_cuda(offset, mask, weight, bias)
    if groups != 1:
        raise NotImplementedError("groups != 1 is not used by GLARE (deformableDecoder_arch.py:151-152)")
    x, offset, mask = f32c(input), f32c(offset), f32c(mask)
    B, C, H, W = x.shape
    Co, Cw, kh, kw = weight.shape
    if Cw != C:
        raise ValueError("weight expects %d input channels, input has %d" % (Cw, C))
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    if tuple(offset.shape) != (B, deformable_groups * 2 * kh * kw, Ho, Wo):
        raise ValueError("offset shape %s != %s" % (tuple(offset.shape), (B, deformable_groups * 2 * kh * kw, Ho, Wo)))
    if tuple(mask.shape) != (B, deformable_groups * kh * kw, Ho, Wo):
        raise ValueError("mask shape %s != %s" % (tuple(mask.shape), (B, deformable_groups * kh * kw, Ho, Wo)))
    if (packed_weight is None and (kh, kw, stride, padding, dilation) == (3, 3, 1, 1, 1) and C % 32 == 0 and (C // deformable_groups) % 8 == 0
            and Co % 4 == 0 and deformable_groups * 27 <= 108):
        # GLARE's configuration (deformableDecoder_arch.py:129-131: 3x3, stride 1, pad 1, dg 4, 128 / 256 channels): the tensor-core kernel
        # (csrc/dcn_tc.cu), 3-5 x the reference extension's speed on B200 (profiles/r43_dcn_ref_compare.txt); other shapes: the fp32 FMA kernel
        return _modulated_deform_conv_tc(x, offset, mask, weight, bias, deformable_groups)
    if packed_weight is None:
        packed_weight = dcn_pack_weight(weight)
    y = x.new_empty((B, Co, Ho, Wo))
    b = f32c(bias) if bias is not None else None
    check(lib().glare_dcnv2_fwd_f32(ptr(x), ptr(offset), ptr(mask), ptr(packed_weight), ptr(b), B, C, H, W, Co, kh, kw,
                                    stride, padding, dilation, deformable_groups, ptr(y), stream()), "glare_dcnv2_fwd_f32")
    return y


_DCN_W = {}


def _modulated_deform_conv_tc(x, offset, mask, weight, bias, dg):
    """NCHW operator inputs -> NHWC, offset | mask side by side, fp32-grade (bf16x3) tensor-core kernel, NCHW-shaped (channels_last) result"""
    import weakref
    B, C, H, W = x.shape
    Co = weight.shape[0]
    key = (weight.data_ptr(), tuple(weight.shape))
    ent = _DCN_W.get(key)
    if ent is None or ent[0]() is not weight or ent[1] != weight._version:
        if len(_DCN_W) > 64:
            _DCN_W.clear()
        ent = _DCN_W[key] = (weakref.ref(weight), weight._version, conv_pack_weight(MODE_BF16X3, weight)[0])
    xn = x.permute(0, 2, 3, 1).contiguous()
    om = torch.cat((offset.permute(0, 2, 3, 1), mask.permute(0, 2, 3, 1)), dim=3).contiguous()
    y = torch.empty((B, H, W, Co), device=x.device, dtype=torch.float32)
    b = f32c(bias) if bias is not None else None
    check(lib().glare_dcnv2_fwd_nhwc_tc(MODE_BF16X3, ptr(xn), ptr(om), ptr(ent[2]), None, ptr(b), ptr(y), B, H, W, C, Co, dg, stream()),
          "glare_dcnv2_fwd_nhwc_tc")
    return y.permute(0, 3, 1, 2)


def dcnv2_bwd_data(x_nhwc, offset, mask, dcol_nhwc, dg, grad_x_nhwc, grad_offset, grad_mask, col_nhwc):
    """deform_conv_cuda.cpp:571-685 data half: grad input (accumulated), grad offset, grad mask, im2col operand"""
    require_cuda(x_nhwc, offset, mask, dcol_nhwc, grad_x_nhwc, grad_offset, grad_mask, col_nhwc)
    B, H, W, C = x_nhwc.shape
    check(lib().glare_dcnv2_bwd_data_f32(ptr(x_nhwc), ptr(offset), ptr(mask), ptr(dcol_nhwc), B, C, H, W, dg, ptr(grad_x_nhwc), ptr(grad_offset),
                                         ptr(grad_mask), ptr(col_nhwc), stream()), "glare_dcnv2_bwd_data_f32")


def dcnv2_bwd_weight(col_nhwc, gout_nhwc, grad_w_packed):
    """grad_w_packed [9C][Cout] += col^T gout over all pixels of the batch"""
    require_cuda(col_nhwc, gout_nhwc, grad_w_packed)
    P = col_nhwc.shape[0] * col_nhwc.shape[1] * col_nhwc.shape[2]
    check(lib().glare_dcnv2_bwd_weight_f32(ptr(col_nhwc), ptr(gout_nhwc), P, col_nhwc.shape[3], gout_nhwc.shape[3], ptr(grad_w_packed),
                                           stream()), "glare_dcnv2_bwd_weight_f32")


def wgrad_conv_tc(x_nhwc, gy_nhwc, k, stride, pad, chunk=8192, mode=None, min_rows=128):
    """weight gradient of a k x k conv, [k*k*Ci][Co] (tap-major rows), on the tensor cores WITHOUT fp32 im2col columns: the transposed bf16x3
    operands of im2col(x) and of dY are written directly (csrc/train_wgrad.cu), then one batched tcgen05 GEMM over the pixel chunks, summed
    in fp32.  Co is zero-padded to a multiple of 32.  None when the shape is outside that path (Ci % 32, fewer than 128 rows).
    mode: MODE_BF16X3 (default, fp32-grade) or MODE_BF16 (single-piece bf16 operands: the bf16 training configuration)."""
    import torch.nn.functional as F
    mode = MODE_BF16X3 if mode is None else mode
    if mode not in (MODE_BF16, MODE_BF16X3):
        return None
    q, e2 = (32, 2) if mode == MODE_BF16X3 else (64, 1)                 # K granule in pixels, operand entries per pixel
    build = lib().glare_im2col_t_operand_bf16x3 if mode == MODE_BF16X3 else lib().glare_im2col_t_operand_bf16
    B, H, W, Ci = x_nhwc.shape
    _, Ho, Wo, Co = gy_nhwc.shape
    M, P = k * k * Ci, B * Ho * Wo
    if M < min_rows or M % 16 or Ci % 32 or not x_nhwc.is_contiguous() or not gy_nhwc.is_contiguous() or x_nhwc.dtype != torch.float32:
        return None
    N = (Co + 31) // 32 * 32
    if N != Co:
        gy_nhwc = F.pad(gy_nhwc, (0, N - Co))
    chunk = max(q, min(chunk, (P + q - 1) // q * q) // q * q)
    nch = (P + chunk - 1) // chunk
    a_op = torch.empty((nch, M, e2 * chunk), device=x_nhwc.device, dtype=torch.bfloat16)
    b_op = torch.empty((nch, N, e2 * chunk), device=x_nhwc.device, dtype=torch.bfloat16)
    check(build(ptr(x_nhwc), B, H, W, Ci, k, stride, pad, Ho, Wo, chunk, ptr(a_op), stream()), "glare_im2col_t_operand")
    check(build(ptr(gy_nhwc), B, Ho, Wo, N, 1, 1, 0, Ho, Wo, chunk, ptr(b_op), stream()), "glare_im2col_t_operand")
    y = torch.empty((nch, M, N), device=x_nhwc.device, dtype=torch.float32)
    conv2d_nhwc_tc_ex(mode, a_op, None, b_op, None, y, nch, M // 16, 16, chunk, N, N, N * chunk)
    return y[:, :, :Co].sum(dim=0)


# ------------------------------------------------------------------------------------------- flow
def flow_cond_tail(p, p_strides, nets, n_steps, nout, B, h, w, out, out_batch_stride, out_step_stride):
    """p_strides = (batch, step, channel, pixel) element strides of the pre-activation planes"""
    require_cuda(p, nets, out)
    check(lib().glare_flow_cond_tail_f32(ptr(p), p_strides[0], p_strides[1], p_strides[2], p_strides[3], ptr(nets), n_steps, nout,
                                         B, h, w, ptr(out), out_batch_stride, out_step_stride, stream()), "glare_flow_cond_tail_f32")


def flow_step(direction, coupling, z_in, z_out, pA, pA_strides, hF, hF_batch_stride, netA, pw, logdet=None):
    """pA_strides = (batch, channel, pixel) element strides"""
    require_cuda(z_in, z_out, pA, hF, netA, pw, logdet)
    B, _, h, w = z_in.shape
    check(lib().glare_flow_step_f32(direction, 1 if coupling else 0, ptr(z_in), ptr(z_out), ptr(pA), pA_strides[0], pA_strides[1],
                                    pA_strides[2], ptr(hF), hF_batch_stride, ptr(netA), ptr(pw), B, h, w, ptr(logdet), stream()),
          "glare_flow_step_f32")


# ------------------------------------------------------------------------------------------- dense convs / GroupNorm
MODE_BF16, MODE_TF32, MODE_TF32X3, MODE_TF32_BF16X2, MODE_BF16X3 = 0, 1, 2, 3, 4


def _hi_alloc(mode, shape, device):
    """primary operand tensor: bf16 (mode 0), fp32 (modes 1-3), interleaved bf16 pair with 2 entries per element (mode 4)"""
    if mode == MODE_BF16:
        return torch.empty(shape, device=device, dtype=torch.bfloat16)
    if mode == MODE_BF16X3:
        return torch.empty(tuple(shape[:-1]) + (2 * shape[-1],), device=device, dtype=torch.bfloat16)
    return torch.empty(shape, device=device, dtype=torch.float32)


def _lo_like(mode, hi):
    """second operand tensor of the fp32-grade modes: fp32 lo (mode 2) or the interleaved bf16 x tensor, 2 per element (mode 3)"""
    if mode == MODE_TF32X3:
        return torch.empty_like(hi)
    if mode == MODE_TF32_BF16X2:
        return torch.empty(tuple(hi.shape[:-1]) + (2 * hi.shape[-1],), device=hi.device, dtype=torch.bfloat16)
    return None


def conv_pack_weight(mode, w_oihw):
    """OIHW fp32 -> (hi, lo|None) packed [Cout][kh*kw][Cin] operands for glare_conv2d_nhwc_tc"""
    require_cuda(w_oihw)
    w = f32c(w_oihw)
    Co, Ci, kh, kw = w.shape
    hi = _hi_alloc(mode, (Co, kh * kw, Ci), w.device)
    lo = _lo_like(mode, hi)
    check(lib().glare_conv_pack_weight(mode, ptr(w), Co, Ci, kh, ptr(hi), ptr(lo), stream()), "glare_conv_pack_weight")
    return hi, lo


def conv_prep_act(mode, x_nhwc):
    """fp32 NHWC activations -> (hi, lo|None) operands (mode 1 passes the tensor through)"""
    require_cuda(x_nhwc)
    if mode == MODE_TF32:
        return x_nhwc, None
    hi = _hi_alloc(mode, tuple(x_nhwc.shape), x_nhwc.device)
    lo = _lo_like(mode, hi)
    check(lib().glare_conv_prep_act(mode, ptr(x_nhwc), x_nhwc.numel(), ptr(hi), ptr(lo), stream()), "glare_conv_prep_act")
    return hi, lo


def conv2d_nhwc_tc(mode, x_hi, x_lo, w_hi, w_lo, bias, residual, B, H, W, Cin, Cout, ksize, out=None):
    """x_* NHWC [B,H,W,Cin] operands, w_* packed; returns y NHWC [B,H,W,Cout] fp32"""
    require_cuda(x_hi, x_lo, w_hi, w_lo, bias, residual)
    y = out if out is not None else torch.empty((B, H, W, Cout), device=x_hi.device, dtype=torch.float32)
    check(lib().glare_conv2d_nhwc_tc(mode, ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(bias), ptr(residual), ptr(y), B, H, W, Cin,
                                     Cout, ksize, stream()), "glare_conv2d_nhwc_tc")
    return y


def dcnv2_pack_fwd_nhwc_tc(mode, x_nhwc, offmask_nhwc, w_hi, w_lo, bias, B, H, W, C, Cout, dg):
    """DCNv2Pack tail (deformableDecoder_arch.py:141-152) on tensor cores; returns y NHWC [B,H,W,Cout] fp32"""
    require_cuda(x_nhwc, offmask_nhwc, w_hi, w_lo, bias)
    y = torch.empty((B, H, W, Cout), device=x_nhwc.device, dtype=torch.float32)
    check(lib().glare_dcnv2_pack_fwd_nhwc_tc(mode, ptr(x_nhwc), ptr(offmask_nhwc), ptr(w_hi), ptr(w_lo), ptr(bias), ptr(y), B, H, W, C,
                                             Cout, dg, stream()), "glare_dcnv2_pack_fwd_nhwc_tc")
    return y


def conv2d_nhwc_tc_g(mode, kind, x_hi, x_lo, w_hi, w_lo, bias, residual, y, B, Hin, Win, Cin, Cout, ksize=3, pa=0, pb=0, gn_stats=None):
    """general conv entry (kind 0 stride-1, 1 Downsample, 2 Upsample phase).  gn_stats ([B,32,2] fp64, kinds 0 / 1): GroupNorm statistics of
    the output from the epilogue (per-tile partials in a scratch tensor + a small fp64 finish launch inside the same call)"""
    global LAUNCHES
    require_cuda(x_hi, x_lo, w_hi, w_lo, bias, residual, y, gn_stats)
    scratch, n = None, 0
    if gn_stats is not None:
        n = int(lib().glare_conv_gn_scratch_floats(B, y.shape[1], y.shape[2]))
        scratch = torch.empty((n,), device=y.device, dtype=torch.float32)
        LAUNCHES += 1                                   # the finish kernel
    check(lib().glare_conv2d_nhwc_tc_g(mode, kind, ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(bias), ptr(residual), ptr(y), B, Hin, Win,
                                       Cin, Cout, ksize, pa, pb, ptr(gn_stats), ptr(scratch), n, stream()), "glare_conv2d_nhwc_tc_g")
    return y


def conv2d_nhwc_tc_down2(mode, x_hi, x_lo, w_hi, w_lo, bias, B, Hin, Win, Cin, Cout):
    """Downsample.forward (encoder_decoder.py:68-72) -> y NHWC [B,Ho,Wo,Cout] fp32"""
    require_cuda(x_hi, x_lo, w_hi, w_lo, bias)
    Ho, Wo = (Hin - 2) // 2 + 1, (Win - 2) // 2 + 1
    y = torch.empty((B, Ho, Wo, Cout), device=x_hi.device, dtype=torch.float32)
    check(lib().glare_conv2d_nhwc_tc_down2(mode, ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(bias), ptr(y), B, Hin, Win, Cin, Cout,
                                           stream()), "glare_conv2d_nhwc_tc_down2")
    return y


def conv2d_nhwc_tc_up2(mode, x_hi, x_lo, phase_weights, bias, B, H, W, Cin, Cout):
    """Upsample.forward (encoder_decoder.py:49-53) as four sub-pixel phase convolutions -> y NHWC [B,2H,2W,Cout] fp32.
    phase_weights[(a, b)] = (w_hi, w_lo) packed 2x2 filters"""
    require_cuda(x_hi, x_lo, bias)
    y = torch.empty((B, 2 * H, 2 * W, Cout), device=x_hi.device, dtype=torch.float32)
    for (a, b), (w_hi, w_lo) in phase_weights.items():
        check(lib().glare_conv2d_nhwc_tc_up2_phase(mode, ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(bias), ptr(y), B, H, W, Cin, Cout,
                                                   a, b, stream()), "glare_conv2d_nhwc_tc_up2_phase")
    return y


def conv2d_nhwc_tc_ex(mode, x_hi, x_lo, w_hi, w_lo, y, B, H, W, Cin, Cout, ldy, w_batch_stride, bias=None, ksize=1):
    """form with an output pixel stride and per-sample weights (attention GEMMs, Cout % 4 != 0 heads); writes into ``y``"""
    require_cuda(x_hi, x_lo, w_hi, w_lo, y, bias)
    check(lib().glare_conv2d_nhwc_tc_ex(mode, ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), ptr(bias), None, ptr(y), B, H, W, Cin, Cout,
                                        ksize, ldy, w_batch_stride, stream()), "glare_conv2d_nhwc_tc_ex")
    return y


def attn_softmax_rows(mode, S, rows, lds, n_keys, n_pad, scale, out_hi, out_lo, ldp):
    require_cuda(S, out_hi, out_lo)
    check(lib().glare_attn_softmax_rows(mode, ptr(S), rows, lds, n_keys, n_pad, scale, ptr(out_hi), ptr(out_lo), ldp, stream()),
          "glare_attn_softmax_rows")


def attn_transpose_v(mode, v_nhwc, B, N, C, Np):
    require_cuda(v_nhwc)
    hi = _hi_alloc(mode, (B, C, Np), v_nhwc.device)
    lo = _lo_like(mode, hi)
    check(lib().glare_attn_transpose_v(mode, ptr(v_nhwc), B, N, C, Np, ptr(hi), ptr(lo), stream()), "glare_attn_transpose_v")
    return hi, lo


def attn_row_norm(x_rows, rows, C, rows_per_sample, norm_out=None, max_bits=None):
    """|x_row| per row and / or the per-sample maximum (bits of a non-negative float, zeroed by the caller)"""
    require_cuda(x_rows, norm_out, max_bits)
    check(lib().glare_attn_row_norm(ptr(x_rows), rows, C, rows_per_sample, ptr(norm_out), ptr(max_bits), stream()), "glare_attn_row_norm")


def attn_row_ref(s_sub, rows, lds, n_sub, scale, offset, ref_out):
    """per-row softmax reference from the scores against a sampled subset of the keys: scale * rowmax + offset"""
    require_cuda(s_sub, ref_out)
    check(lib().glare_attn_row_ref(ptr(s_sub), rows, lds, n_sub, scale, offset, ptr(ref_out), stream()), "glare_attn_row_ref")


def attn_scores_exp_tc(mode, q_op, k_op, rows_h, rows_w, C, n_keys, n_pad, scale, margin, q_norm, key_max, p_out, row_sum_part, part_stride):
    """scores GEMM with exp(scale * s - ref(row)) in the epilogue; returns the number of output blocks (valid planes of row_sum_part)"""
    import ctypes
    require_cuda(q_op, k_op, q_norm, key_max, p_out, row_sum_part)
    nb = ctypes.c_int(0)
    check(lib().glare_attn_scores_exp_tc(mode, ptr(q_op), ptr(k_op), rows_h, rows_w, C, n_keys, n_pad, scale, margin, ptr(q_norm), ptr(key_max),
                                         ptr(p_out), ptr(row_sum_part), part_stride, ctypes.cast(ctypes.byref(nb), ctypes.c_void_p), stream()),
          "glare_attn_scores_exp_tc")
    return nb.value


def attn_row_sum_finish(part, part_stride, n_blocks, rows, row_scale, flag):
    require_cuda(part, row_scale, flag)
    check(lib().glare_attn_row_sum_finish(ptr(part), part_stride, n_blocks, rows, ptr(row_scale), ptr(flag), stream()), "glare_attn_row_sum_finish")


def attn_pv_tc(mode, p_op, vt_op, row_scale, y, rows_h, rows_w, n_pad, C, ldy, pack_out=False, key_band=0):
    """y: fp32 [rows][ldy], or with pack_out the bf16x3 operand [rows][2 * C] of the following conv.  key_band > 0: the contraction over the
    n_pad keys runs in bands of key_band keys chained through the epilogue's residual input (fp32 running sum, round-to-nearest adds), which
    bounds the tensor-core accumulator's truncation bias for very long rows (1080p); the row scale / operand packing ride on the last band."""
    require_cuda(p_op, vt_op, row_scale, y)
    e2 = 2 if mode == MODE_BF16X3 else 1               # operand entries per logical element
    es = p_op.element_size()
    if key_band <= 0 or key_band >= n_pad:
        check(lib().glare_attn_pv_tc(mode, ptr(p_op), 0, ptr(vt_op), 0, ptr(row_scale), None, ptr(y), rows_h, rows_w, n_pad, C, ldy,
                                     1 if pack_out else 0, stream()), "glare_attn_pv_tc")
        return
    import ctypes
    rows = rows_h * rows_w
    acc = [torch.empty((rows, C), device=p_op.device, dtype=torch.float32) for _ in range(2)]
    prev = None
    for i, k0 in enumerate(range(0, n_pad, key_band)):
        nk = min(key_band, n_pad - k0)
        last = k0 + nk >= n_pad
        pp = ctypes.c_void_p(p_op.data_ptr() + k0 * e2 * es)
        vv = ctypes.c_void_p(vt_op.data_ptr() + k0 * e2 * es)
        dst = y if last else acc[i & 1]
        check(lib().glare_attn_pv_tc(mode, pp, n_pad, vv, n_pad, ptr(row_scale) if last else None, ptr(prev), ptr(dst), rows_h, rows_w, nk, C,
                                     ldy if last else C, 1 if (pack_out and last) else 0, stream()), "glare_attn_pv_tc")
        prev = dst


def conv2d_nhwc_tc_pack(mode, x_op, w_op, bias, B, H, W, Cin, Cout, ksize, row_sq=False):
    """stride-1 conv whose output is written as the operand of the next GEMM; returns (operand [B,H,W,2*Cout] bf16, (part, n_blocks) | None)"""
    import ctypes
    require_cuda(x_op, w_op, bias)
    y = _hi_alloc(mode, (B, H, W, Cout), x_op.device)
    part, nb = None, ctypes.c_int(0)
    if row_sq:
        part = torch.empty(((Cout + 63) // 64, B * H * W), device=x_op.device, dtype=torch.float32)
    check(lib().glare_conv2d_nhwc_tc_pack(mode, ptr(x_op), ptr(w_op), ptr(bias), ptr(y), B, H, W, Cin, Cout, ksize, ptr(part), B * H * W,
                                          ctypes.cast(ctypes.byref(nb), ctypes.c_void_p), stream()), "glare_conv2d_nhwc_tc_pack")
    return y, ((part, nb.value) if row_sq else None)


def attn_row_norm_finish(part, n_blocks, rows, rows_per_sample, norm_out=None, max_bits=None):
    require_cuda(part, norm_out, max_bits)
    check(lib().glare_attn_row_norm_finish(ptr(part), part.shape[1], n_blocks, rows, rows_per_sample, ptr(norm_out), ptr(max_bits), stream()),
          "glare_attn_row_norm_finish")


def gn_stats(x_nhwc, B, HW, C, G=32):
    require_cuda(x_nhwc)
    stats = torch.empty((B, G, 2), device=x_nhwc.device, dtype=torch.float64)
    check(lib().glare_gn_stats_nhwc_f32(ptr(x_nhwc), B, HW, C, G, ptr(stats), stream()), "glare_gn_stats_nhwc_f32")
    return stats


def gn_apply(out_mode, x_nhwc, stats, gamma, beta, swish, B, HW, C, G=32, eps=1e-6):
    """-> (hi, lo|None) with hi bf16 (mode 0) / fp32 (mode 1) / tf32-hi (mode 2), same NHWC shape as x"""
    require_cuda(x_nhwc, stats, gamma, beta)
    hi = _hi_alloc(out_mode, tuple(x_nhwc.shape), x_nhwc.device)
    lo = _lo_like(out_mode, hi)
    check(lib().glare_gn_apply_nhwc(out_mode, ptr(x_nhwc), ptr(stats), ptr(gamma), ptr(beta), eps, 1 if swish else 0, B, HW, C, G,
                                    ptr(hi), ptr(lo), stream()), "glare_gn_apply_nhwc")
    return hi, lo


# ------------------------------------------------------------------------------------------- AFT decoder glue
def aft_axpby(a, b, alpha, beta):
    """out = a * alpha + b * beta with alpha / beta device tensors of 1 or B elements (Mix.forward, deformableDecoder_arch.py:587-590, and
    the mean-ratio residual :567); a, b logical [B,C,H,W] fp32 sharing one dense storage order (NHWC or NCHW); rounding as the reference's
    separate mul / mul / add kernels"""
    require_cuda(a, b, alpha, beta)
    if a.shape != b.shape or a.dtype != torch.float32 or b.dtype != torch.float32 or a.stride() != b.stride():
        raise ValueError("aft_axpby needs two fp32 tensors of one shape and storage order")
    if not (a.is_contiguous() or a.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("aft_axpby needs dense NCHW or NHWC storage")
    B = a.shape[0]
    n = a[0].numel()
    alpha, beta = f32c(alpha).reshape(-1), f32c(beta).reshape(-1)
    if alpha.numel() not in (1, B) or beta.numel() not in (1, B) or n % 4:
        raise ValueError("alpha / beta must hold 1 or B values and the sample size must be a multiple of 4")
    out = torch.empty_like(a)
    check(lib().glare_aft_axpby_f32(ptr(a), ptr(b), ptr(alpha), ptr(beta), int(alpha.numel() == B and B > 1), int(beta.numel() == B and B > 1),
                                    B, n, ptr(out), stream()), "glare_aft_axpby_f32")
    return out


def aft_cat_operand(a_nhwc, b_nhwc):
    """bf16x3 operand of cat([a, b], channel) from two NHWC fp32 tensors of the same pixels (no fp32 cat tensor)"""
    require_cuda(a_nhwc, b_nhwc)
    B, H, W, Ca = a_nhwc.shape
    Cb = b_nhwc.shape[3]
    out = torch.empty((B, H, W, 2 * (Ca + Cb)), device=a_nhwc.device, dtype=torch.bfloat16)
    check(lib().glare_aft_cat_operand(ptr(a_nhwc), ptr(b_nhwc), B * H * W, Ca, Cb, ptr(out), stream()), "glare_aft_cat_operand")
    return out


# ------------------------------------------------------------------------------------------- pre / post-processing
def preprocess_u8(img_u8_nhwc, pad, mode):
    """uint8 [B,H,W,3] (device) -> log(clamp(x/255 + 1e-3)) fp32 [B,3,Hp,Wp]; pad = (top, bottom, left, right); mode 0 reflect, 1 symmetric"""
    require_cuda(img_u8_nhwc)
    B, H, W, _ = img_u8_nhwc.shape
    t, b, l, r = pad
    out = torch.empty((B, 3, H + t + b, W + l + r), device=img_u8_nhwc.device, dtype=torch.float32)
    check(lib().glare_preprocess_u8(ptr(img_u8_nhwc), B, H, W, t, b, l, r, mode, ptr(out), stream()), "glare_preprocess_u8")
    return out


def postprocess_u8(y, box):
    """fp32 logical [B,3,Hp,Wp] (any strides) -> uint8 [B,H,W,3] of box = (y0, y1, x0, x1): clip to [0,1], * 255, truncate"""
    require_cuda(y)
    y0, y1, x0, x1 = box
    B = y.shape[0]
    out = torch.empty((B, y1 - y0, x1 - x0, 3), device=y.device, dtype=torch.uint8)
    sb, sc, sh, sw = y.stride()
    check(lib().glare_postprocess_u8(ptr(y), sb, sc, sh, sw, B, y0, x0, y1 - y0, x1 - x0, ptr(out), stream()), "glare_postprocess_u8")
    return out
