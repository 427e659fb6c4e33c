"""CPU: the stage-2 training step of glare_b200/encoder_train.py (tape over ConEncoder1 + the flow objective of flow_train.py).
  * block-level backward rules and the whole step, with torch restatements of the kernel-level primitives, against torch autograd of the
    oracle for EVERY parameter and against the reference's own gradients (tests/golden/stage2.npz);
  * the new kernels of csrc/train_enc.cu: the very source nvcc compiles, executed on the host (tests/cuda_emu), against those restatements."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT, load_golden
from encoder_train_emu import TorchLeaves
from flow_train_emu import TorchEmuKernels


def _step(sd, lr, gt, leaves):
    from glare_b200 import encoder_train, flow
    plan = flow.FlowPlan(sd, torch.device("cpu"))
    conv = lambda x, wgt: F.conv2d(x, wgt, None, padding=1)               # noqa: E731
    with torch.no_grad():
        return encoder_train.stage2_step(sd, plan, lr, gt, leaves, conv, flow_kernels=TorchEmuKernels(sd))


def test_stage2_step_matches_autograd_and_reference():
    from glare_b200 import synth
    from oracle import glare_oracle as O
    g = load_golden("stage2")
    sd = synth.synth_state_dict("netG_stage2", 0)
    gt, lr = torch.from_numpy(g["gt_latent"]), torch.from_numpy(g["lr"])
    nll, grads = _step(sd, lr, gt, TorchLeaves())
    assert np.allclose(nll.numpy(), g["nll"], atol=1e-4, rtol=1e-5)
    # the reference's own gradients (flow and encoder parameters)
    for key in list(g):
        if key.startswith("grad."):
            ref = torch.from_numpy(g[key])
            assert float((grads[key[5:]] - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), key
    # every parameter against torch autograd of the oracle
    sda = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    _, nll_a = O.stage2_nll(sda, gt, lr)
    nll_a.mean().backward()
    checked = 0
    for k, v in sda.items():
        if v.grad is None:
            assert k not in grads or float(grads[k].abs().max()) == 0.0, k
            continue
        assert k in grads, k
        assert grads[k].shape == v.shape, k
        sc = max(float(v.grad.abs().max()), 1e-6)
        assert float((grads[k] - v.grad).abs().max()) <= 1e-3 * sc + 1e-7, (k, float((grads[k] - v.grad).abs().max()), sc)
        checked += 1
    assert checked > 600


def _host_lib():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    emu = os.path.join(ROOT, "tests", "cuda_emu")
    out = os.path.join(emu, "_build", "libtrain_enc_emu.so")
    src = os.path.join(ROOT, "glare_b200", "csrc", "train_enc.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(emu, "cuda_emu.h"))):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-x", "c++", "-DGLARE_CUDA_EMU", "-I", emu, "-shared", "-fPIC", "-pthread", src, "-o", out])
    return ctypes.CDLL(out)


def test_train_enc_kernel_source_on_the_host():
    from glare_b200 import _lib, encoder_train
    host = _host_lib()

    class HostLeaves(encoder_train.CudaLeaves):
        def __init__(self):
            pass

        def _call(self, name, *args):
            fn = getattr(host, name)
            fn.argtypes, fn.restype = _lib.SIGNATURES[name], ctypes.c_int
            assert fn(*args, None) == 0, name

    H, T = HostLeaves(), TorchLeaves()
    gen = torch.Generator().manual_seed(8)

    def close(a, b, what, tol=2e-5):
        sc = max(float(b.abs().max()), 1e-6)
        assert a.shape == b.shape and float((a - b).abs().max()) <= tol * sc, (what, float((a - b).abs().max()), sc)

    # GroupNorm (+ swish) backward
    for C, swish, hw in ((64, True, (3, 4)), (128, False, (3, 4)), (128, True, (9, 23)), (256, True, (5, 7)), (512, False, (2, 3))):
        x = torch.randn((2,) + hw + (C,), generator=gen) * 1.5 + 0.7
        gy = torch.randn((2,) + hw + (C,), generator=gen)
        gamma, beta = 1 + 0.2 * torch.randn(C, generator=gen), 0.1 * torch.randn(C, generator=gen)
        _, stats = T.gn_fwd(x, gamma, beta, swish)
        for nm, a, b in zip(("gx", "dgamma", "dbeta"), H.gn_bwd(x, gy, stats, gamma, beta, swish), T.gn_bwd(x, gy, stats, gamma, beta, swish)):
            close(a, b, "gn_bwd %s C=%d" % (nm, C))
    # im2col for every conv geometry of the encoder
    for (Hh, Ww, C, k, stride, pad) in ((5, 6, 8, 3, 1, 1), (5, 6, 8, 1, 1, 0), (6, 8, 4, 3, 2, 0), (7, 5, 4, 3, 2, 0)):
        x = torch.randn((2, Hh, Ww, C), generator=gen)
        Ho, Wo = (Hh, Ww) if stride == 1 else ((Hh + 1 - 3) // 2 + 1, (Ww + 1 - 3) // 2 + 1)
        close(H.im2col(x, k, stride, pad, Ho, Wo), T.im2col(x, k, stride, pad, Ho, Wo), "im2col %s" % ((Hh, Ww, C, k, stride, pad),), 0.0)
    # bias gradients: column sums of any width that is a multiple of 4, ragged row counts
    for (Pp, C) in ((1000, 128), (37, 64), (513, 516), (5, 4)):
        xs = torch.randn((Pp, C), generator=gen)
        close(H.colsum(xs), xs.double().sum(dim=0).float(), "colsum %dx%d" % (Pp, C), 1e-5)
    # softmax backward
    P = torch.softmax(torch.randn((9, 20), generator=gen), dim=1)
    dP = torch.randn((9, 20), generator=gen)
    close(H.softmax_bwd(P, dP, 0.3), T.softmax_bwd(P, dP, 0.3), "softmax_bwd")


def test_cuda_leaves_glue_with_fake_ops():
    """the padding / leading-dimension logic of encoder_train.CudaLeaves.gemm_nt and .softmax_rows, driven through stand-ins for the two library
    calls they wrap that enforce the C entry points' argument checks (csrc/conv_tc.cu conv_tc_launch, csrc/attn.cu glare_attn_softmax_rows)"""
    from glare_b200 import encoder_train

    class FakeOps:
        @staticmethod
        def conv_prep_act(mode, x):
            assert mode == 1 and x.is_contiguous()
            return x, None

        @staticmethod
        def conv2d_nhwc_tc_ex(mode, x_hi, x_lo, w_hi, w_lo, y, B, H, W, Cin, Cout, ldy, w_batch_stride, bias=None, ksize=1):
            assert Cin % 32 == 0 and ldy >= Cout and ldy % 4 == 0 and not (Cout % 4 != 0 and ldy == Cout)
            assert w_batch_stride in (0, Cout * Cin) and x_hi.numel() == B * H * W * Cin and y.numel() == B * H * W * ldy
            if w_batch_stride == 0:
                assert w_hi.numel() == Cout * Cin
                y.view(-1, ldy)[:, :Cout] = x_hi.reshape(-1, Cin) @ w_hi.reshape(Cout, Cin).t()
            else:                                   # per-sample weights: sample n uses w + n * w_batch_stride
                assert w_hi.numel() == B * Cout * Cin
                y.view(B, H * W, ldy)[:, :, :Cout] = torch.bmm(x_hi.reshape(B, H * W, Cin), w_hi.reshape(B, Cout, Cin).transpose(1, 2))

        @staticmethod
        def attn_softmax_rows(mode, S, rows, lds, n_keys, n_pad, scale, out_hi, out_lo, ldp):
            assert mode == 1 and n_pad >= n_keys and n_pad % 4 == 0 and lds % 4 == 0 and ldp % 4 == 0 and lds >= n_keys and ldp >= n_pad and scale > 0
            assert S.is_contiguous() and S.numel() == rows * lds and out_hi.numel() == rows * ldp
            out_hi.view(rows, ldp).zero_()
            out_hi.view(rows, ldp)[:, :n_keys] = torch.softmax(S.view(rows, lds)[:, :n_keys] * scale, dim=1)

    class FakeDense:
        mode = 1

    L = encoder_train.CudaLeaves.__new__(encoder_train.CudaLeaves)
    L.dense, L.ops = FakeDense(), FakeOps()
    gen = torch.Generator().manual_seed(3)
    for (h, w, K, N) in ((4, 5, 20, 20), (4, 5, 64, 7), (3, 3, 9, 64), (8, 8, 512, 64)):
        a, b = torch.randn((h * w, K), generator=gen), torch.randn((N, K), generator=gen)
        got = L.gemm_nt(a, b, (h, w))
        assert got.shape == (h * w, N) and torch.allclose(got, a @ b.t(), atol=1e-5)
        got = L.gemm_nt_ta(a.t().contiguous(), b, (h, w))                  # modes without a transposed-operand kernel: transposed copy + gemm_nt
        assert got.shape == (h * w, N) and torch.allclose(got, a @ b.t(), atol=1e-5)
    for (P, M, N, chunk) in ((200, 128, 32, 64), (97, 144, 40, 8192), (64, 9, 64, 32)):
        a, b = torch.randn((P, M), generator=gen), torch.randn((P, N), generator=gen)
        got = L.gemm_tn_tc(a, b, chunk=chunk)
        assert got.shape == (M, N) and torch.allclose(got, a.t() @ b, atol=1e-4)
    for (R, N) in ((6, 20), (5, 7), (3, 64)):
        S = torch.randn((R, N), generator=gen)
        got = L.softmax_rows(S, 0.25)
        assert got.shape == (R, N) and got.is_contiguous() and torch.allclose(got, torch.softmax(S * 0.25, dim=1), atol=1e-6)


def test_stage2_step_non_square_single_sample():
    """a second geometry (48 x 32 input, latent 12 x 8, batch 1): odd tile counts in every level, against autograd of the oracle"""
    from glare_b200 import synth
    from oracle import glare_oracle as O
    sd = synth.synth_state_dict("netG_stage2", 0)
    gen = torch.Generator().manual_seed(12)
    gt = torch.randn((1, 3, 12, 8), generator=gen)
    lr = synth.preprocess(torch.rand((1, 3, 48, 32), generator=gen))
    nll, grads = _step(sd, lr, gt, TorchLeaves())
    sda = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    _, nll_a = O.stage2_nll(sda, gt, lr)
    nll_a.mean().backward()
    assert torch.allclose(nll, nll_a.detach(), atol=1e-4, rtol=1e-5)
    bad = []
    for k, v in sda.items():
        if v.grad is not None:
            sc = max(float(v.grad.abs().max()), 1e-6)
            if float((grads[k] - v.grad).abs().max()) > 1e-3 * sc + 1e-7:
                bad.append((k, float((grads[k] - v.grad).abs().max()) / sc))
    assert len(bad) <= 2 and all(e < 5e-2 for _, e in bad), bad      # isolated ReLU sign flips of the split first-layer sum (see flow_train_gpu_check.py)


def test_stage2_nll_autograd_node_fills_param_grads_like_the_reference():
    """encoder_train.stage2_nll over a module's named parameters: after ``.mean().backward()`` every ``param.grad`` equals the reference's
    (tests/golden/stage2.npz), including under a loss-scaling factor"""
    from glare_b200 import encoder_train, synth
    g = load_golden("stage2")
    params = {k: torch.nn.Parameter(v.clone()) for k, v in synth.synth_state_dict("netG_stage2", 0).items()}
    gt, lr = torch.from_numpy(g["gt_latent"]), torch.from_numpy(g["lr"])
    conv = lambda x, wgt: F.conv2d(x, wgt, None, padding=1)               # noqa: E731
    sd_val = {k: p.detach() for k, p in params.items()}
    nll = encoder_train.stage2_nll(params.items(), gt, lr, TorchLeaves(), conv, flow_kernels=TorchEmuKernels(sd_val))
    assert np.allclose(nll.detach().numpy(), g["nll"], atol=1e-4, rtol=1e-5)
    (nll.mean() * 128.0).backward()                                       # GradScaler-style factor
    for key in list(g):
        if key.startswith("grad."):
            ref, got = torch.from_numpy(g[key]) * 128.0, params[key[5:]].grad
            assert got is not None and float((got - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), key
    assert params["flowUpsamplerNet.layers.0.actnorm.bias"].grad is not None


def test_stage2_step_gt_mean_branch_matches_reference():
    """the `mean = gt` branch of LLFlowVQGAN2.normal_flow (LLFlowVQGAN2_arch.py:109, probability train_gt_ratio = 0.2): objective and gradients
    against the reference's own (tests/golden/stage2_gtmean.npz) and against autograd of the oracle; color_conv receives no gradient"""
    from glare_b200 import encoder_train, flow, synth
    from oracle import glare_oracle as O
    g, gm = load_golden("stage2"), load_golden("stage2_gtmean")
    sd = synth.synth_state_dict("netG_stage2", 0)
    gt, lr = torch.from_numpy(g["gt_latent"]), torch.from_numpy(g["lr"])
    conv = lambda x, wgt: F.conv2d(x, wgt, None, padding=1)               # noqa: E731
    with torch.no_grad():
        nll, grads = encoder_train.stage2_step(sd, flow.FlowPlan(sd, torch.device("cpu")), lr, gt, TorchLeaves(), conv,
                                               flow_kernels=TorchEmuKernels(sd), use_gt_mean=True)
    assert np.allclose(nll.numpy(), gm["nll"], atol=1e-4, rtol=1e-5)
    assert not np.allclose(gm["nll"], g["nll"], atol=1e-2)                # really the other branch
    for key in list(gm):
        if key.startswith("grad."):
            ref = torch.from_numpy(gm[key])
            assert float((grads[key[5:]] - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), key
    for k in gm["no_grad"]:
        assert str(k) not in grads
    sda = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    _, nll_a = O.stage2_nll(sda, gt, lr, use_gt_mean=True)
    nll_a.mean().backward()
    for k, v in sda.items():
        if v.grad is None:
            assert k not in grads or float(grads[k].abs().max()) == 0.0, k
        else:
            sc = max(float(v.grad.abs().max()), 1e-6)
            assert float((grads[k] - v.grad).abs().max()) <= 1e-3 * sc + 1e-7, k


def test_draw_use_gt_mean_follows_the_reference_rng():
    import random
    from glare_b200 import encoder_train
    random.seed(1)
    want = [not (random.random() > 0.2) for _ in range(50)]
    random.seed(1)
    assert [encoder_train.draw_use_gt_mean(0.2) for _ in range(50)] == want and any(want) and not all(want)


def test_packed_weight_cache_does_not_grow_with_optimizer_steps():
    """the cache was keyed on the weight's in-place version -> one more packed copy of every weight per training step"""
    from glare_b200.dense import TcDense
    d = TcDense.__new__(TcDense)
    d._w, d.mode, d.force_repack, calls = {}, 4, False, []

    class Ops:
        @staticmethod
        def conv_pack_weight(mode, w):
            calls.append(w._version)
            return (w.clone(), None)

    d.ops = Ops()
    w = torch.zeros((8, 32, 3, 3))
    for step in range(5):
        d._weights(w)
        d._weights(w)                      # second use inside the step: cached
        w.add_(1.0)                        # optimizer step bumps the version
    assert len(calls) == 5 and len(d._w) == 1
    for _ in range(600):                   # per-step temporaries (new tensor objects) are swept once dead
        d._weights(torch.zeros((8, 32, 1, 1)))
    assert len(d._w) <= 257
