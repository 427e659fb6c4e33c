"""CPU oracle for the GLARE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package -- as the checker, never as the thing shipped.  ``glare_b200`` (the
product) must not import it.

Parity status: PINNED against the reference's own Python modules imported from ``/root/reference``
in the authoring container (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``); the reference ships
no tests/golden vectors of its own.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "oracle_kernels.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        f32p = ctypes.POINTER(ctypes.c_float)
        _LIB.glare_oracle_vq_f32.argtypes = [f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_int64), f32p, f32p]
        _LIB.glare_oracle_vq_f32.restype = None
        _LIB.glare_oracle_dcn_im2col_f32.argtypes = [f32p, f32p, f32p] + [ctypes.c_int] * 10 + [f32p]
        _LIB.glare_oracle_dcn_im2col_f32.restype = None
    return _LIB


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def vq_lookup(z_nchw, codebook, want_dmin=False):
    """z [B,3,h,w] fp32, codebook [K,3] fp32 -> (idx int64 [B*h*w], z_q [B,3,h,w]).  quantize.py:271-312."""
    z, zp = _f32(z_nchw)
    cb, cbp = _f32(codebook)
    B, C, h, w = z.shape
    assert C == 3 and cb.shape[1] == 3
    idx = np.empty(B * h * w, dtype=np.int64)
    zq = np.empty_like(z)
    dmin = np.empty(B * h * w, dtype=np.float32) if want_dmin else None
    lib().glare_oracle_vq_f32(zp, cbp, B, h * w, cb.shape[0],
                              idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                              zq.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                              dmin.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if want_dmin else None)
    return (idx, zq, dmin) if want_dmin else (idx, zq)


def dcn_im2col(x, offset, mask, kh=3, kw=3, stride=1, pad=1, dil=1, dg=4):
    """Literal restatement of modulated_deformable_im2col (deform_conv_cuda_kernel.cu:571-633).
    Returns cols [B, C*kh*kw, Ho*Wo]."""
    x, xp = _f32(x)
    offset, op = _f32(offset)
    mask, mp = _f32(mask)
    B, C, H, W = x.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    assert offset.shape == (B, dg * 2 * kh * kw, Ho, Wo) and mask.shape == (B, dg * kh * kw, Ho, Wo)
    cols = np.empty((B, C * kh * kw, Ho * Wo), dtype=np.float32)
    lib().glare_oracle_dcn_im2col_f32(xp, op, mp, B, C, H, W, kh, kw, stride, pad, dil, dg,
                                      cols.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return cols
