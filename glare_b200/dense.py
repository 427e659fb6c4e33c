"""Dense-operator backends of the engine and their roofline bookkeeping (see engine.py docstring).

``make_dense("torch-fp32")``  cuDNN/cuBLAS fp32 (TF32 off) -- library baseline, bring-up backend
``make_dense("torch-bf16")``  cuDNN/cuBLAS bf16            -- library baseline
``make_dense("auto")``        the best hand-written backend available in the built library, else torch-fp32
"""
import torch

from .engine import TorchDense


def dcn_flops(C, Cout, px):
    return 2.0 * C * Cout * 9 * px


class _TorchBackend(TorchDense):
    def __init__(self, dtype, allow_tf32=False):
        super().__init__(dtype, allow_tf32)
        self.dtype_name = {torch.float32: "fp32", torch.bfloat16: "bf16"}[dtype]
        self.name = "torch-library-" + self.dtype_name

    def roofline(self, timers, eng, B, lr_shape, pk):
        """Dominant kernel of libglare_b200.so in this configuration: the DCNv2 forward (fp32 FMA path).
        Algorithmic work = 2*C*Cout*9 FLOP per output pixel: scale 0 C=256 @ H/2 x W/2,
        scale 1 C=128 @ H x W."""
        Hp, Wp = lr_shape[2], lr_shape[3]
        out = {}
        tot_ms, tot_fl, n = 0.0, 0.0, 0
        for i, (C, px) in enumerate(((256, (Hp // 2) * (Wp // 2)), (128, Hp * Wp))):
            ev = timers.get("dcn%d" % i, [])
            if not ev:
                continue
            ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            tot_ms += ms
            tot_fl += dcn_flops(C, C, px) * B
            n += 1
        if not n:
            return None
        ach = tot_fl / (tot_ms / 1e3) / 1e12
        return {"kernel": "dcn_fwd_kernel (both AFT scales, fp32 FMA)", "bound": "tensor", "achieved": ach, "peak": pk["tensor"],
                "unit": "TFLOP/s", "frac": ach / pk["tensor"], "traffic": None, "ms_per_step": tot_ms, "peak_source": pk["src"]}


def make_dense(name="auto"):
    if name in ("auto", "torch-fp32"):
        return _TorchBackend(torch.float32)
    if name == "torch-tf32":
        b = _TorchBackend(torch.float32, allow_tf32=True)
        b.name, b.dtype_name = "torch-library-tf32", "tf32"
        return b
    if name == "torch-bf16":
        return _TorchBackend(torch.bfloat16)
    raise ValueError("unknown dense backend %r" % name)
