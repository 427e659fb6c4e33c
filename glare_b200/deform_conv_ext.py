"""Drop-in for the reference's native plug-in module ``deform_conv_ext`` (pybind11, ops/dcn/src/deform_conv_ext.cpp:150-164), same five
function names and argument lists, kernels from libglare_b200.so.  The reference binds it with ``from . import deform_conv_ext``
(ops/dcn/deform_conv.py:23-26); placing (or symlinking) this file at ``code/models/modules/ops/dcn/deform_conv_ext.py`` -- or registering it
as ``sys.modules['models.modules.ops.dcn.deform_conv_ext']`` before the import, see INTEGRATION.md -- makes the reference's own
``ModulatedDeformConvFunction`` (deform_conv.py:121-184) run on these kernels with no other rebinding.

Contract kept from the reference: the CALLER allocates ``output`` and every gradient tensor (zero-initialised,
deform_conv.py:147,161-165) and passes two empty scratch tensors that stay untouched; inputs are contiguous NCHW / OIHW CUDA tensors;
nothing is retained across calls; the kernels go to the current CUDA stream; shape errors raise RuntimeError, CPU tensors
NotImplementedError.  Only the two ``modulated_*`` functions are reached by GLARE; the plain (v1) deformable convolution is not implemented.
"""
import torch

from . import ops


def _check(input, weight, offset, mask, group, deformable_group, kernel_h, kernel_w):
    for t in (input, weight, offset, mask):
        if not t.is_cuda:
            raise NotImplementedError("deform_conv_ext is not implemented on CPU")            # deform_conv_cuda.cpp AT_ERROR on CPU tensors
    if not (input.is_contiguous() and weight.is_contiguous()):
        raise RuntimeError("input / weight tensor has to be contiguous")                        # deform_conv_cuda.cpp:497-498
    if group != 1:
        raise RuntimeError("glare_b200 deform_conv_ext: group != 1 is not used by GLARE (deformableDecoder_arch.py:151-152)")
    if tuple(weight.shape[2:]) != (kernel_h, kernel_w) or weight.shape[1] != input.shape[1]:
        raise RuntimeError("Input shape and kernel shape wont match: (%d x %d vs %d x %d)" % (kernel_h, kernel_w, weight.shape[2], weight.shape[3]))
    if input.shape[1] % deformable_group:
        raise RuntimeError("input channels must be divisible by deformable_group")


def modulated_deform_conv_forward(input, weight, bias, ones, offset, mask, output, columns, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                  pad_w, dilation_h, dilation_w, group, deformable_group, with_bias):
    """deform_conv_cuda.cpp:490-569 modulated_deform_conv_cuda_forward: writes ``output`` [B,Cout,Ho,Wo] in place"""
    _check(input, weight, offset, mask, group, deformable_group, kernel_h, kernel_w)
    if stride_h != stride_w or pad_h != pad_w or dilation_h != dilation_w:
        raise RuntimeError("glare_b200 deform_conv_ext: stride / padding / dilation must be equal in h and w")
    y = ops.modulated_deform_conv(input, offset, mask, weight, bias if with_bias else None, stride_h, pad_h, dilation_h, group, deformable_group)
    if tuple(output.shape) != tuple(y.shape):
        raise RuntimeError("output shape %s, expected %s" % (tuple(output.shape), tuple(y.shape)))
    output.copy_(y)


def modulated_deform_conv_backward(input, weight, bias, ones, offset, mask, columns, grad_input, grad_weight, grad_bias, grad_offset, grad_mask,
                                   grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, group,
                                   deformable_group, with_bias):
    """deform_conv_cuda.cpp:571-685 modulated_deform_conv_cuda_backward: ACCUMULATES into the caller's (zeroed) gradient tensors like the
    reference's per-sample loop does"""
    from .dcn_backward import dcn_backward
    _check(input, weight, offset, mask, group, deformable_group, kernel_h, kernel_w)
    if (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w) != (3, 3, 1, 1, 1, 1, 1, 1):
        raise RuntimeError("glare_b200 DCN backward covers GLARE's configuration: 3x3, stride 1, pad 1, dilation 1")
    gi, go, gm, gw, gb = dcn_backward(input, offset, mask, weight, grad_output, deformable_group, bool(with_bias))
    grad_input.add_(gi)
    grad_offset.add_(go)
    grad_mask.add_(gm)
    grad_weight.add_(gw)
    if with_bias:
        grad_bias.add_(gb)


def _v1(*args, **kwargs):
    raise NotImplementedError("glare_b200 deform_conv_ext: the unmodulated (v1) deformable convolution is not reached by any GLARE configuration "
                              "(only ModulatedDeformConvFunction is, deformableDecoder_arch.py:151-152)")


deform_conv_forward = _v1                 # deform_conv_ext.cpp:52-75
deform_conv_backward_input = _v1          # :77-100
deform_conv_backward_parameters = _v1     # :102-123
