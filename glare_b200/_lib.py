"""ctypes binding of libglare_b200.so (include/glare_b200.h).  Fails loudly: no fallback of any kind."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglare_b200.so")
_LIB = None

_vp, _i, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES); the single source the CPU test compares with the header
SIGNATURES = {
    "glare_abi_version": [],
    "glare_error_string": [_i],
    "glare_vq_pack_codebook_f32": [_vp, _i, _vp, _vp],
    "glare_vq_argmin_gather_f32": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "glare_flow_net_floats": [],
    "glare_flow_cond_tail_f32": [_vp, _ll, _ll, _ll, _ll, _vp, _i, _i, _i, _i, _i, _vp, _ll, _ll, _vp],
    "glare_flow_step_f32": [_i, _i, _vp, _vp, _vp, _ll, _ll, _ll, _vp, _ll, _vp, _vp, _i, _i, _i, _vp, _vp],
    "glare_flow_train_net_fwd_f32": [_vp, _ll, _vp, _ll, _vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "glare_flow_train_point_fwd_f32": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "glare_flow_train_coupling_bwd_f32": [_i, _vp, _vp, _vp, _vp, ctypes.c_float, _i, _i, _i, _vp, _vp, _vp],
    "glare_flow_train_net_bwd_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _vp, _vp],
    "glare_flow_train_point_bwd_f32": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "glare_flow_train_im2col3x3_f32": [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp],
    "glare_gemm_tn_skinny_f32": [_vp, _vp, _ll, _i, _i, _vp, _vp],
    "glare_flow_train_colsum_f32": [_vp, _ll, _vp, _ll, _i, _ll, _vp, _vp],
    "glare_dcn_pack_weight_f32": [_vp, _i, _i, _i, _i, _vp, _vp],
    "glare_dcnv2_fwd_f32": [_vp, _vp, _vp, _vp, _vp] + [_i] * 11 + [_vp, _vp],
    "glare_dcnv2_bwd_data_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "glare_dcnv2_bwd_weight_f32": [_vp, _vp, _ll, _i, _i, _vp, _vp],
    "glare_dcnv2_pack_fwd_nhwc_tc": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "glare_dcnv2_fwd_nhwc_tc": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "glare_conv_tc_elem_bytes": [_i],
    "glare_conv_pack_weight": [_i, _vp, _i, _i, _i, _vp, _vp, _vp],
    "glare_conv_prep_act": [_i, _vp, _ll, _vp, _vp, _vp],
    "glare_conv2d_nhwc_tc": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "glare_conv2d_nhwc_tc_down2": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "glare_conv2d_nhwc_tc_up2_phase": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "glare_conv2d_nhwc_tc_g": [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _ll, _vp],
    "glare_conv_gn_scratch_floats": [_i, _i, _i],
    "glare_conv2d_nhwc_tc_ex": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _ll, _ll, _vp],
    "glare_attn_softmax_rows": [_i, _vp, _ll, _ll, _i, _i, ctypes.c_float, _vp, _vp, _ll, _vp],
    "glare_attn_transpose_v": [_i, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "glare_attn_row_norm": [_vp, _ll, _i, _ll, _vp, _vp, _vp],
    "glare_attn_row_ref": [_vp, _ll, _ll, _i, ctypes.c_float, ctypes.c_float, _vp, _vp],
    "glare_attn_scores_exp_tc": [_i, _vp, _vp, _i, _i, _i, _i, _i, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _ll, _vp, _vp],
    "glare_attn_row_sum_finish": [_vp, _ll, _i, _ll, _vp, _vp, _vp],
    "glare_attn_pv_tc": [_i, _vp, _ll, _vp, _ll, _vp, _vp, _vp, _i, _i, _i, _i, _ll, _i, _vp],
    "glare_conv2d_nhwc_tc_pack": [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _vp],
    "glare_attn_row_norm_finish": [_vp, _ll, _i, _ll, _ll, _vp, _vp, _vp],
    "glare_gn_bwd_nhwc_f32": [_vp, _vp, _vp, _vp, _vp, ctypes.c_float, _i, _i, _ll, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "glare_im2col_nhwc_f32": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "glare_im2col_t_operand_bf16x3": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "glare_im2col_t_operand_bf16": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "glare_colsum_f32": [_vp, _ll, _i, _vp, _vp],
    "glare_attn_softmax_bwd_f32": [_vp, _vp, _ll, _ll, _i, ctypes.c_float, _vp, _vp],
    "glare_ssim_partials": [_i, _i, _i, _i],
    "glare_ssim_fwd_f32": [_vp, _vp, _i, _i, _i, _i, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp],
    "glare_ssim_bwd_f32": [_vp, _vp, _i, _i, _i, _i, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "glare_relu_f32": [_vp, _vp, _ll, _vp, _vp],
    "glare_maxpool2_nhwc_f32": [_vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "glare_maxpool2_nhwc_bwd_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "glare_avgpool2_f32": [_vp, _ll, _i, _i, _vp, _vp],
    "glare_up2_nhwc_f32": [_vp, _i, _i, _i, _i, _i, _vp, _vp],
    "glare_aft_axpby_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _ll, _vp, _vp],
    "glare_aft_cat_operand": [_vp, _vp, _ll, _i, _i, _vp, _vp],
    "glare_preprocess_u8": [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "glare_postprocess_u8": [_vp, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _vp, _vp],
    "glare_gn_stats_nhwc_f32": [_vp, _i, _ll, _i, _i, _vp, _vp],
    "glare_gn_apply_nhwc": [_i, _vp, _vp, _vp, _vp, ctypes.c_float, _i, _i, _ll, _i, _i, _vp, _vp, _vp],
}
_RESTYPES = {"glare_error_string": ctypes.c_char_p, "glare_conv_gn_scratch_floats": ctypes.c_longlong,
             "glare_ssim_partials": ctypes.c_longlong}


class GlareLibraryError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GlareLibraryError(
                "libglare_b200.so is not built (%s). Run `python -m glare_b200.build`; glare_b200 has no "
                "fallback path." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library does not export a declared symbol
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _LIB = L
    return _LIB


def check(code, what):
    if code != 0:
        msg = lib().glare_error_string(int(code))
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", code))


def stream():
    """The caller's current CUDA stream, as the reference op uses at::cuda::getCurrentCUDAStream()."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            # the reference op raises NotImplementedError on CPU tensors (ops/dcn/deform_conv.py:143-144)
            raise NotImplementedError("glare_b200 operators run on CUDA (sm_100a) tensors only; got a %s tensor" % t.device)


def f32c(t):
    """contiguous fp32 view/copy (the reference TORCH_CHECKs contiguity; callers here get a copy instead)"""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
