"""TEST INFRASTRUCTURE: cuDNN / cuBLAS implementation of the engine's dense-backend interface (conv2d, gn_swish, attention), fp32 with TF32
off by default.  It is the on-device comparison path of the GPU tests and probes (the tcgen05 kernels against the vendor library on the same
inputs) and is never imported by the package: glare_b200 has no library backend."""
import torch
import torch.nn.functional as F


class TorchDense:
    """Library (cuDNN/cuBLAS) implementation of the dense operators, fp32 or bf16.  This is the baseline the
    hand-written tensor-core path is measured against, and the dense backend of the first bring-up."""
    name = "torch-library"

    def __init__(self, dtype=torch.float32, allow_tf32=False):
        self.dtype = dtype
        self.allow_tf32 = allow_tf32

    def _ctx(self):
        torch.backends.cudnn.allow_tf32 = self.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.allow_tf32

    def conv2d(self, x, w, b=None, stride=1, padding=1, residual=None):
        self._ctx()
        y = F.conv2d(x.to(self.dtype), w.to(self.dtype), None if b is None else b.to(self.dtype), stride=stride,
                     padding=padding)
        return y if residual is None else y + residual

    def gn_swish(self, x, gamma, beta, swish=True):
        y = F.group_norm(x.float(), 32, gamma, beta, eps=1e-6)       # encoder_decoder.py:34-35
        if swish:
            y = y * torch.sigmoid(y)                                 # encoder_decoder.py:29-31
        return y.to(self.dtype)

    def attention(self, q, k, v):
        """single-head attention over h*w tokens, d = C (encoder_decoder.py:176-187); q,k,v [B,C,h,w]"""
        self._ctx()
        B, C, h, w = q.shape
        out = torch.empty_like(q)
        for b in range(B):                       # the N x N score matrix is materialised per sample (1 GB at 105x155)
            qb = q[b].reshape(C, h * w).t()
            s = torch.mm(qb, k[b].reshape(C, h * w)) * (int(C) ** (-0.5))
            s = torch.softmax(s.float(), dim=1).to(q.dtype)
            out[b] = torch.mm(v[b].reshape(C, h * w), s.t()).reshape(C, h, w)
        return out
