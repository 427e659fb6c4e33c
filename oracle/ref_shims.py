"""Import harness for the UNMODIFIED reference (LowLevelAI/GLARE) on CPU.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) validate the oracle restatement in
``oracle/glare_oracle.py`` against the reference's own Python modules and (b) generate the golden
vectors committed under ``tests/golden/`` (see ``oracle/gen_golden.py``).  It only works in the
authoring container where ``/root/reference`` is mounted; nothing on the GPU box imports it.

Shim set: stub third-party modules the reference imports but the hot path never
uses, replace the VGG feature extractors that download weights at construction, ignore the
hard-coded ``'cuda'`` device literals when no GPU is present, and route the CUDA-only DCN op
through ``torchvision.ops.deform_conv2d`` (same channel layout and border rule as
``ops/dcn/src/deform_conv_cuda_kernel.cu:571-633``).
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("GLARE_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")


def available():
    return os.path.isdir(REF_CODE)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


_installed = False


def install():
    """Idempotently prepare ``sys.modules`` / ``sys.path`` so ``import models...`` resolves to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REF_ROOT)
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)

    _stub("pytorch_lightning", LightningModule=nn.Module)
    _stub("lpips", LPIPS=lambda *a, **k: nn.Identity())
    _stub("natsort", natsorted=sorted, natsort=types.SimpleNamespace(natsorted=sorted))    # `from natsort import natsort` (infer_unpaired.py:5)
    _stub("pyiqa")
    _stub("tensorboardX", SummaryWriter=object)
    sk = _stub("skimage")
    skm = _stub("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)
    sk.metrics = skm

    # hard-coded device literals -> no-ops on a CPU-only host
    if not torch.cuda.is_available():
        _to = torch.Tensor.to

        def to(self, *a, **k):
            a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
            if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
                k.pop("device")
            return _to(self, *a, **k) if (a or k) else self

        torch.Tensor.to = to
        torch.Tensor.cuda = lambda self, *a, **k: self
        _mto = nn.Module.to

        def mto(self, *a, **k):
            a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
            return _mto(self, *a, **k) if (a or k) else self

        nn.Module.to = mto
        nn.Module.cuda = lambda self, *a, **k: self

    # VGG extractors download weights at construction; neither is used by inference or stage 2
    import models.modules.vgg_arch as vgg_arch

    class _NoVGG(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    vgg_arch.VGGFeatureExtractor = _NoVGG
    import models.modules.VQModel_arch as vqm
    vqm.VGGFeatureExtractor = _NoVGG
    import models.modules.losses as losses
    losses.RefPerceptualNetwork = losses.PerceptualNetwork          # kept for oracle/gen_golden_stage3.py (built without its __init__)
    losses.PerceptualNetwork = _NoVGG

    # CUDA-only DCN -> torchvision CPU kernel
    import torchvision.ops
    import models.modules.deformableDecoder_arch as dda

    def _dcn(x, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups):
        assert groups == 1
        return torchvision.ops.deform_conv2d(x, offset, weight, bias, stride=stride, padding=padding,
                                             dilation=dilation, mask=mask)

    dda.modulated_deform_conv = _dcn

    # stage 2 ships with a broken import (LLFlowVQGAN2_arch.py:10)
    import models.modules.ConditionEncoder as ce
    import models.modules.VQGANConditionEncoder as vce
    if not hasattr(ce, "NoEncoder"):
        ce.NoEncoder = vce.NoEncoder
    _installed = True


def parse_opt(yml="LOL.yml"):
    install()
    import options.options as option
    cwd = os.getcwd()
    try:
        opt = option.parse(os.path.join(REF_CODE, "confs", yml), is_train=False)
    finally:
        os.chdir(cwd)
    return option.dict_to_nonedict(opt)


def build_reference(yml="LOL.yml", seed=0):
    """Returns (netG, net_hq) = the reference's VQLLFLOWDeformable generator and VQModel, on CPU, eval mode."""
    install()
    import models.networks as networks
    opt = parse_opt(yml)
    torch.manual_seed(seed)
    import numpy as np
    np.random.seed(seed)
    netG = networks.define_Flow(opt, 0).eval()
    net_hq = networks.find_vqgan(opt).eval()
    return netG, net_hq, opt


def build_reference_stage2(seed=0):
    """Stage-2 generator (LLFlowVQGAN2) + VQModel (train_stage2_LOL.yml)."""
    install()
    import models.networks as networks
    opt = parse_opt("train_stage2_LOL.yml")
    torch.manual_seed(seed)
    import numpy as np
    np.random.seed(seed)
    netG = networks.define_Flow(opt, 0)
    net_hq = networks.find_vqgan(opt).eval()
    return netG, net_hq, opt
