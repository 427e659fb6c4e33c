"""CPU: the arithmetic behind the fused-softmax attention (glare_b200/csrc/attn.cu, conv_tc.cu epilogues) restated in numpy --
softmax against a Cauchy-Schwarz row reference instead of the row maximum, two-piece bf16 operands, 1/rowsum applied after the P V
product -- against the reference formulation (encoder_decoder.py:176-187) in fp64.  Specification-level checks; the kernels themselves
are compared with fp64 on the GPU (tests/test_conv_tc_gpu.py)."""
import numpy as np
import torch


def _bf16(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).bfloat16().float().numpy()


def _split_b3(x):
    a1 = _bf16(x)
    return a1, _bf16(x - a1)


def _fused_attention(q, k, v, margin=60.0):
    """q, k, v [N, C] fp32 -> (out [N, C], row sums) the way the kernels evaluate it"""
    C = q.shape[1]
    scale = np.float32(C ** -0.5)
    qn = np.sqrt((q.astype(np.float32) ** 2).sum(1)) * np.float32(1.00001)
    kmax = (np.sqrt((k.astype(np.float32) ** 2).sum(1)) * np.float32(1.00001)).max()
    ref = qn * (scale * kmax) - np.float32(margin)                       # >= every logit of the row - margin
    s = (q.astype(np.float64) @ k.astype(np.float64).T).astype(np.float32)
    t = s * (scale * np.float32(1.4426950408889634)) - (ref * np.float32(1.4426950408889634))[:, None]
    p = np.exp2(t.astype(np.float32)).astype(np.float32)                 # ex2 of the fp32 exponent
    rowsum = p.sum(1, dtype=np.float32)
    p1, p2 = _split_b3(p)
    v1, v2 = _split_b3(v)
    acc = p1.astype(np.float64) @ v1 + p1.astype(np.float64) @ v2 + p2.astype(np.float64) @ v1   # three bf16 passes, wide accumulator
    return (acc / rowsum[:, None]).astype(np.float32), rowsum


def _reference_attention(q, k, v):
    C = q.shape[1]
    w = torch.softmax(torch.from_numpy(q).double() @ torch.from_numpy(k).double().t() * (int(C) ** -0.5), dim=1)
    return (w @ torch.from_numpy(v).double()).numpy()


def test_row_reference_is_shift_invariant_and_fp32_grade():
    rng = np.random.default_rng(0)
    N, C = 300, 512
    q = (rng.standard_normal((N, C)) * 2.0).astype(np.float32)
    k = (rng.standard_normal((N, C)) * 1.5).astype(np.float32)
    v = rng.standard_normal((N, C)).astype(np.float32)
    out, rowsum = _fused_attention(q, k, v)
    ref = _reference_attention(q, k, v)
    assert np.abs(out - ref).max() < 3e-5 * max(1.0, np.abs(ref).max())
    assert np.all(rowsum >= 1e-24) and np.all(np.isfinite(rowsum))       # nothing would raise the device flag


def test_window_violation_is_detectable():
    """|q| |k| C^-0.5 far above the true logits: every exponent underflows, the row sum leaves [1e-24, 3e38] -> flag"""
    rng = np.random.default_rng(1)
    N, C = 64, 512
    q = (rng.standard_normal((N, C)) * 0.05).astype(np.float32)
    k = (rng.standard_normal((N, C)) * 0.05).astype(np.float32)
    q[:, 0], k[:, 0], k[:, 1], q[:, 1] = 400.0, 0.0, 400.0, 0.0
    v = rng.standard_normal((N, C)).astype(np.float32)
    with np.errstate(all="ignore"):
        _, rowsum = _fused_attention(q, k, v)
    assert not np.all((rowsum >= 1e-24) & (rowsum <= 3e38))


def test_bf16x3_operand_layout_roundtrip():
    """operand tensors of mode 4: per 32-element K chunk [32 x a1 | 32 x a2] (common.cuh store_b3_4); Operand.dense() inverts it"""
    from glare_b200.dense import Operand
    rng = np.random.default_rng(2)
    B, H, W, C = 1, 3, 5, 64
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    a1, a2 = _split_b3(x)
    inter = np.stack([a1.reshape(-1, 32), a2.reshape(-1, 32)], axis=1).reshape(B, H, W, 2 * C)
    op = Operand(4, torch.from_numpy(inter).bfloat16(), None, B, C, H, W)
    back = op.dense().permute(0, 2, 3, 1).numpy()
    assert np.abs(back - x).max() <= 2.0 ** -16 * np.abs(x).max()        # two bf16 pieces: 16 significant bits
    assert np.array_equal(back, a1 + a2)
