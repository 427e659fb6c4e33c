"""ModulatedDeformConvFunction with a backward pass (ops/dcn/deform_conv.py:121-184), kernels from libglare_b200.so.

forward : the fp32 operator (glare_dcnv2_fwd_f32) -- same semantics and argument order as the reference Function.
backward: for the configuration GLARE trains with (3x3, stride 1, pad 1, dilation 1, groups 1; stage 3, VQLLFLOWD_model.py:187-232):
    dcol   = W^T grad_out            1x1 conv on the tcgen05 path (fp32-grade mode)
    grad_input / grad_offset / grad_mask / col   one fused gather-scatter kernel (glare_dcnv2_bwd_data_f32)
    grad_weight = col^T grad_out     batched tcgen05 GEMM over pixel chunks (ops.wgrad_conv_tc; fp32 split-K GEMM when 9C % 32 != 0)
    grad_bias   = sum grad_out
Unlike the reference there is no per-sample host loop and no `columns` buffer shared between calls; `chunk` bounds the two
[n,H,W,9C] scratch tensors (1.2 GB per sample at 128 channels, 420x620).
"""
import os

import torch

from . import ops


class ModulatedDeformConvFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1):
        if not input.is_cuda:
            raise NotImplementedError                                     # deform_conv.py:143-144
        ctx.cfg = (stride, padding, dilation, groups, deformable_groups)
        ctx.with_bias = bias is not None
        if weight.requires_grad or mask.requires_grad or offset.requires_grad or input.requires_grad:
            ctx.save_for_backward(input, offset, mask, weight)
        return ops.modulated_deform_conv(input.detach(), offset.detach(), mask.detach(), weight.detach(),
                                         None if bias is None else bias.detach(), stride, padding, dilation, groups, deformable_groups)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        stride, padding, dilation, groups, dg = ctx.cfg
        input, offset, mask, weight = ctx.saved_tensors
        Co, C, kh, kw = weight.shape
        if (kh, kw, stride, padding, dilation, groups) != (3, 3, 1, 1, 1, 1):
            raise NotImplementedError("glare_b200 DCN backward covers GLARE's configuration: 3x3, stride 1, pad 1, dilation 1, groups 1")
        gi, go, gm, gw, gb = dcn_backward(input, offset, mask, weight, grad_output, dg, ctx.with_bias)
        return gi, go, gm, gw, gb, None, None, None, None, None


modulated_deform_conv = ModulatedDeformConvFunction.apply
_MODE = ops.MODE_TF32_BF16X2                     # fp32-grade tensor-core mode for dcol = W^T grad_out
_WGRAD_TC = not os.environ.get("GLARE_WGRAD_FMA")   # GLARE_WGRAD_FMA=1: the fp32 split-K GEMM (glare_dcnv2_bwd_weight_f32), the correctness baseline


def dcn_backward(x, offset, mask, weight, grad_output, dg, with_bias=True, chunk=2):
    """returns (grad_input, grad_offset, grad_mask, grad_weight, grad_bias) in the reference's layouts (all NCHW / OIHW)"""
    B, C, H, W = x.shape
    Co = weight.shape[0]
    if Co % 32 or (C // dg) % 4:
        raise NotImplementedError("DCN backward needs Cout % 32 == 0 and (C / deformable_groups) % 4 == 0 (GLARE: 128 / 256 channels)")
    x_n = x.detach().float().permute(0, 2, 3, 1).contiguous()
    g_n = grad_output.detach().float().permute(0, 2, 3, 1).contiguous()
    offset, mask = offset.detach().float().contiguous(), mask.detach().float().contiguous()
    # W^T as a 1x1 filter with 9C output channels (t*C + c) over Co input channels
    w_t = weight.detach().float().permute(2, 3, 1, 0).reshape(9 * C, Co, 1, 1).contiguous()
    w_hi, w_lo = ops.conv_pack_weight(_MODE, w_t)
    grad_x = torch.zeros_like(x_n)
    grad_offset, grad_mask = torch.empty_like(offset), torch.empty_like(mask)
    grad_wp = torch.zeros((9 * C, Co), device=x.device, dtype=torch.float32)
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        gc = g_n[b0:b1]
        g_hi, g_lo = ops.conv_prep_act(_MODE, gc)
        dcol = ops.conv2d_nhwc_tc(_MODE, g_hi, g_lo, w_hi, w_lo, None, None, b1 - b0, H, W, Co, 9 * C, 1)
        col = torch.empty_like(dcol)
        ops.dcnv2_bwd_data(x_n[b0:b1], offset[b0:b1], mask[b0:b1], dcol, dg, grad_x[b0:b1], grad_offset[b0:b1], grad_mask[b0:b1], col)
        gw = ops.wgrad_conv_tc(col, gc, 1, 1, 0) if _WGRAD_TC else None       # col^T grad_out on the tensor cores (fp32-grade bf16x3)
        if gw is not None:
            grad_wp += gw
        else:
            ops.dcnv2_bwd_weight(col, gc, grad_wp)
    grad_w = grad_wp.view(3, 3, C, Co).permute(3, 2, 0, 1).contiguous()
    grad_b = g_n.sum(dim=(0, 1, 2)) if with_bias else None
    return grad_x.permute(0, 3, 1, 2), grad_offset, grad_mask, grad_w, grad_b
