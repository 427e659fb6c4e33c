"""CPU: host side of the stage-2 flow training step (glare_b200/flow_train.py: 28-step loops, offsets into the hoisted pre-activation
tensor, packed-parameter layout, assembly of the gradients into state-dict shapes) driven through the torch restatement of the kernel
contracts (tests/flow_train_emu.py), against the specification oracle/flow_backward.py -- which tests/test_oracle.py ties to autograd and
to the reference's own gradients."""
import torch
import torch.nn.functional as F

from flow_train_emu import TorchEmuKernels


def test_flow_training_step_host_logic_matches_specification(sd_g):
    from glare_b200 import flow, flow_train
    from oracle import flow_backward as FB
    gen = torch.Generator().manual_seed(21)
    B, h, w = 2, 5, 7
    gt = torch.randn((B, 3, h, w), generator=gen)
    ft = torch.sigmoid(torch.randn((B, 64, h, w), generator=gen))
    mean = torch.randn((B, 3, h, w), generator=gen) * 0.1
    plan = flow.FlowPlan(sd_g, torch.device("cpu"))
    conv = lambda x, wgt: F.conv2d(x, wgt, None, padding=1)               # noqa: E731
    with torch.no_grad():
        nll, z, g_gt, g_ft, g_mean, grads = flow_train.nll_forward_backward(plan, sd_g, gt, ft, mean, conv, kernels=TorchEmuKernels(sd_g))
        nll_s, z_s, g_gt_s, g_ft_s, g_mean_s, grads_s = FB.nll_forward_backward(sd_g, gt, ft, mean)

    def close(a, b, what):
        assert a.shape == b.shape, (what, a.shape, b.shape)
        sc = max(float(b.abs().max()), 1e-6)
        assert float((a - b).abs().max()) <= 2e-4 * sc + 1e-7, (what, float((a - b).abs().max()), sc)

    close(nll, nll_s, "nll"), close(z, z_s, "z"), close(g_gt, g_gt_s, "gt"), close(g_ft, g_ft_s, "ft"), close(g_mean, g_mean_s, "mean")
    assert sorted(grads) == sorted(grads_s)
    for k in grads_s:
        close(grads[k], grads_s[k], k)
        assert tuple(grads[k].shape) == tuple(sd_g[k].shape), k


def test_stage2_step_through_autograd_node_matches_reference_gradients():
    """BASELINE config 4 wiring: cond encoder under torch autograd (the oracle's restatement on the CPU) + the flow objective as the
    FlowNLL autograd node (host logic of glare_b200/flow_train.py on the emulated kernels) against the REFERENCE's own forward and
    backward (tests/golden/stage2.npz from LLFlowVQGAN2_arch.py:75-122): objective, and the gradients of flow and encoder parameters"""
    import numpy as np
    from conftest import load_golden
    from glare_b200 import flow, flow_train, synth
    from oracle import glare_oracle as O
    g = load_golden("stage2")
    sd2 = {k: v.clone().requires_grad_(True) for k, v in synth.synth_state_dict("netG_stage2", 0).items()}
    gt, lr = torch.from_numpy(g["gt_latent"]), torch.from_numpy(g["lr"])
    enc = O.cond_encoder(sd2, lr, "RRDB")                                   # differentiable torch ops
    sd_val = {k: v.detach() for k, v in sd2.items()}
    plan = flow.FlowPlan(sd_val, torch.device("cpu"))
    keys = flow_train.flow_parameter_keys(sd_val)
    conv = lambda x, wgt: F.conv2d(x, wgt, None, padding=1)                # noqa: E731
    nll = flow_train.FlowNLL.apply(plan, sd_val, conv, TorchEmuKernels(sd_val), keys, gt, enc["cond_feat"], enc["color_map"], *[sd2[k] for k in keys])
    assert np.allclose(nll.detach().numpy(), g["nll"], atol=1e-4, rtol=1e-5)
    nll.mean().backward()
    for key in list(g):
        if key.startswith("grad."):
            ref, got = torch.from_numpy(g[key]), sd2[key[5:]].grad
            assert got is not None, key
            assert float((got - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max())), key


# ------------------------------------------------------------------------------------------------------------------------------------
# The ACTUAL kernel source of csrc/flow_bwd.cu, compiled for the host through tests/cuda_emu/cuda_emu.h (threads of a block as OS
# threads, __syncthreads / shuffles / atomics with their CUDA meaning) and driven through the same C ABI.
def _build_host_kernels():
    import ctypes
    import os
    import shutil
    import subprocess
    import pytest
    from conftest import ROOT
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    emu = os.path.join(ROOT, "tests", "cuda_emu")
    out = os.path.join(emu, "_build", "libflow_bwd_emu.so")
    src = os.path.join(ROOT, "glare_b200", "csrc", "flow_bwd.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(emu, "cuda_emu.h"))):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-x", "c++", "-DGLARE_CUDA_EMU", "-I", emu, "-shared", "-fPIC", "-pthread", src, "-o", out])
    return ctypes.CDLL(out)


def test_flow_training_kernel_source_on_the_host(sd_g):
    """every kernel of csrc/flow_bwd.cu (the very source nvcc compiles) executed on the CPU against the torch restatement of its contract, then the
    whole training step through them against the specification -- what tests/test_zz_flow_train_gpu.py repeats on the GPU"""
    import ctypes
    from glare_b200 import _lib, flow_train
    import flow_train_gpu_check as chk
    host = _build_host_kernels()
    emu = TorchEmuKernels(sd_g)

    class HostCompiledKernels(flow_train.CudaKernels):
        def __init__(self):
            pass

        def _call(self, name, *args):
            fn = getattr(host, name)
            fn.argtypes, fn.restype = _lib.SIGNATURES[name], ctypes.c_int
            assert fn(*args, None) == 0, name

        encode_chain = emu.encode_chain          # the forward chain kernels (csrc/flow.cu) are GPU-validated separately
        wgrad = None                             # tensor-core weight gradient (csrc/train_wgrad.cu): GPU only, tests/flow_train_gpu_check.py

        def gemm_tn(self, a, M, b, N, P, out):   # the skinny kernel runs here; the split-K GEMM of csrc/dcn_bwd.cu is GPU-validated (test_dcn_gpu.py)
            if M <= 32 and N <= 256 and 256 % N == 0:
                flow_train.CudaKernels.gemm_tn(self, a, M, b, N, P, out)
            else:
                emu.gemm_tn(a, M, b, N, P, out)

    chk.OK = True
    conv = lambda x, wgt: F.conv2d(x, wgt, None, padding=1)               # noqa: E731
    # small shape: the host shim spawns one OS thread per CUDA thread
    assert chk.run_checks(torch.device("cpu"), HostCompiledKernels(), emu, conv, shape=(2, 4, 5))
