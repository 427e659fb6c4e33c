"""CPU: the stage-3 losses (glare_b200/losses.py: MS-SSIM on csrc/loss.cu, VGG16 perceptual on the conv tape).
  * the oracle restatement (oracle/losses.py) against the reference's own functions (value and gradient);
  * csrc/loss.cu -- the very source nvcc compiles, executed on the host (tests/cuda_emu) -- through MsssimFn against autograd of the oracle;
  * the perceptual loss with torch restatements of the kernel-level primitives against autograd of the oracle."""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch

from conftest import ROOT
from encoder_train_emu import TorchLeaves


def _images(B, H, W, seed=0, noise=0.15):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand((B, 3, H, W), generator=g)
    gt = torch.nn.functional.avg_pool2d(gt, 5, 1, 2)                      # some spatial structure
    sr = (gt + noise * torch.randn((B, 3, H, W), generator=g)).clamp(0, 1)
    return sr, gt


def _vgg_sd(seed=0, scale=1.0):
    from glare_b200 import losses
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for idx, ci, co in losses.VGG_CONVS:
        sd["%d.weight" % idx] = torch.randn((co, ci, 3, 3), generator=g) * (scale * (2.0 / (9 * ci)) ** 0.5)
        sd["%d.bias" % idx] = 0.05 * torch.randn((co,), generator=g)
    return sd


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="reference tree not mounted")
def test_oracle_losses_match_the_reference():
    from oracle import losses as OL
    from oracle import ref_shims
    ref_shims.install()
    from models.modules.pytorch_msssim import msssim as ref_msssim
    import models.modules.losses as ref_losses
    for (B, H, W) in ((2, 176, 192), (1, 40, 36)):
        sr, gt = _images(B, H, W, seed=H)
        for normalize in (True, False):
            a, b = sr.clone().requires_grad_(True), sr.clone().requires_grad_(True)
            va, vb = OL.msssim(a, gt, normalize=normalize), ref_msssim(b, gt, normalize=normalize)
            assert float((va - vb).abs()) <= 1e-6, (H, normalize, float(va), float(vb))
            va.backward()
            vb.backward()
            assert float((a.grad - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max()) + 1e-10
    # the perceptual network: the reference's forward / output_features code over torchvision's vgg16 structure with seeded weights
    from torchvision.models import vgg16
    sd = _vgg_sd(1)
    net = ref_losses.RefPerceptualNetwork.__new__(ref_losses.RefPerceptualNetwork)
    torch.nn.Module.__init__(net)
    net.vgg_model = vgg16(weights=None).features[:16]
    net.vgg_model.load_state_dict(sd, strict=True)
    net.layer_name_mapping = {'3': "relu1_2", '8': "relu2_2", '15': "relu3_3"}
    sr, gt = _images(1, 32, 40, seed=3)
    a, b = sr.clone().requires_grad_(True), sr.clone().requires_grad_(True)
    va, vb = OL.perceptual(sd, a, gt), net(b, gt)
    assert float((va - vb).abs()) <= 1e-6 * float(vb.abs())
    va.backward()
    vb.backward()
    assert float((a.grad - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max())


def _host_lib():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    emu = os.path.join(ROOT, "tests", "cuda_emu")
    out = os.path.join(emu, "_build", "libloss_emu.so")
    src = os.path.join(ROOT, "glare_b200", "csrc", "loss.cu")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(emu, "cuda_emu.h"))):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-x", "c++", "-DGLARE_CUDA_EMU", "-I", emu, "-shared", "-fPIC", "-pthread", src, "-o", out])
    return ctypes.CDLL(out)


def _host_kernels():
    from glare_b200 import _lib, losses
    host = _host_lib()

    class HostSsim(losses.SsimKernels):
        def __init__(self):
            pass

        def _call(self, name, *args):
            fn = getattr(host, name)
            fn.argtypes, fn.restype = _lib.SIGNATURES[name], ctypes.c_int
            assert fn(*args, None) == 0, name

        def _partials(self, planes, H, W, ws):
            fn = host.glare_ssim_partials
            fn.argtypes, fn.restype = _lib.SIGNATURES["glare_ssim_partials"], ctypes.c_longlong
            return int(fn(planes, H, W, ws))

    return HostSsim()


@pytest.mark.parametrize("shape,normalize,val_range", [((1, 40, 36), True, None), ((2, 52, 67), True, 1), ((1, 36, 44), False, None)])
def test_msssim_kernel_source_on_the_host(shape, normalize, val_range):
    """value and d/d(img1) of glare_b200.losses.msssim with csrc/loss.cu run on the host, against autograd of the oracle; the small sizes
    exercise real_size = min(11, H, W) at the coarse levels (:37) and odd sizes under avg_pool2d"""
    from glare_b200 import losses
    from oracle import losses as OL
    K = _host_kernels()
    sr, gt = _images(*shape, seed=7)
    a, b = sr.clone().requires_grad_(True), sr.clone().requires_grad_(True)
    va = losses.msssim(a, gt, normalize=normalize, val_range=val_range, kernels=K)
    vb = OL.msssim(b, gt, normalize=normalize, val_range=val_range)
    assert float((va - vb).abs()) <= 2e-6, (float(va), float(vb))
    (va * 0.7).backward()
    (vb * 0.7).backward()
    assert float((a.grad - b.grad).abs().max()) <= 2e-4 * float(b.grad.abs().max()), float((a.grad - b.grad).abs().max())


def test_elementwise_kernel_source_on_the_host():
    """ReLU / MaxPool2d(2,2) / avg_pool2d / nearest x2 and their backward rules (csrc/loss.cu on the host) against the torch restatements,
    including odd sizes and ties (all-zero windows after a ReLU: the first position wins, like ATen)"""
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
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 8, 7, 10), generator=g)
    x[0, :, 2:4, 4:6] = 0.0                                                # a tie
    gy = torch.randn(x.shape, generator=g)
    y = H.relu(x)
    assert torch.equal(y, T.relu(x)) and torch.equal(H.relu_bwd(y, gy), T.relu_bwd(y, gy))
    for xin in (x, T.relu(x)):
        (yp, saved_h), (yt, saved_t) = H.maxpool2(xin), T.maxpool2(xin)
        assert torch.equal(yp, yt)
        gp = torch.randn(yt.shape, generator=g)
        assert torch.equal(H.maxpool2_bwd(gp, saved_h), T.maxpool2_bwd(gp, saved_t))
    assert torch.allclose(H.avgpool2(x), T.avgpool2(x), atol=1e-7)
    up = H.up2(x)
    assert torch.equal(up, T.up2(x))
    gu = torch.randn(up.shape, generator=g)
    assert torch.allclose(H.up2_adjoint(gu), T.up2_adjoint(gu), atol=1e-6)


def test_msssim_argument_checks():
    from glare_b200 import losses
    sr, gt = _images(1, 40, 36)
    with pytest.raises(NotImplementedError):
        losses.msssim(sr, gt)                                              # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        losses.msssim(sr, gt, size_average=False, kernels=object())
    with pytest.raises(NotImplementedError):
        losses.msssim(sr, gt.clone().requires_grad_(True), kernels=object())


def test_perceptual_matches_autograd_of_the_oracle():
    from glare_b200 import losses
    from oracle import losses as OL
    sd = _vgg_sd(2)
    net = losses.PerceptualNetwork(state_dict={"features." + k: v for k, v in sd.items()}, leaves=TorchLeaves())
    assert sorted(net.state_dict()) == sorted("vgg_model." + k for k in sd)              # the reference's keys (losses.py:15-16)
    assert not any(p.requires_grad for p in net.parameters())                            # :17-18
    sr, gt = _images(2, 24, 32, seed=9)
    a, b = sr.clone().requires_grad_(True), sr.clone().requires_grad_(True)
    va, vb = net(a, gt), OL.perceptual(sd, b, gt)
    assert float((va - vb).abs()) <= 1e-5 * float(vb.abs())
    (va * 0.01).backward()
    (vb * 0.01).backward()
    assert float((a.grad - b.grad).abs().max()) <= 1e-4 * float(b.grad.abs().max())
    feats = net.output_features(sr)
    assert [tuple(f.shape[1:]) for f in feats] == [(64, 24, 32), (128, 12, 16), (256, 6, 8)]
    with pytest.raises(NotImplementedError):
        losses.PerceptualNetwork()(sr, gt)                                                # CPU tensors without injected leaves: no fallback


def test_stage3_objective_matches_the_oracle():
    """|sr - gt| + 0.01 perceptual + 0.2 (1 - MS-SSIM) on a reconstruction with out-of-range and NaN pixels (VQLLFLOWD_model.py:212-223)"""
    from glare_b200 import losses
    from oracle import losses as OL
    sd = _vgg_sd(4)
    net = losses.PerceptualNetwork(state_dict=sd, leaves=TorchLeaves())
    K = _host_kernels()
    sr, gt = _images(1, 48, 40, seed=11, noise=0.4)
    rec = sr + 0.3 * torch.randn(sr.shape, generator=torch.Generator().manual_seed(2))  # values outside [0, 1]
    rec[0, 1, 5, 7] = float("nan")
    a, b = rec.clone().requires_grad_(True), rec.clone().requires_grad_(True)
    ta, parts_a = losses.stage3_loss(a, gt, net, msssim_fn=lambda x, y, **kw: losses.msssim(x, y, kernels=K, **kw))
    tb, parts_b = OL.stage3_loss(b, gt, sd)
    for k in parts_b:
        assert float((parts_a[k] - parts_b[k]).abs()) <= 1e-5 * max(1e-3, float(parts_b[k].abs())), k
    ta.backward()
    tb.backward()
    assert float(a.grad[0, 1, 5, 7]) == 0.0 and float(b.grad[0, 1, 5, 7]) == 0.0
    assert float((a.grad - b.grad).abs().max()) <= 2e-4 * float(b.grad.abs().max())
