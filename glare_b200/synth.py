"""Seeded synthetic checkpoints and inputs for GLARE (no weights or datasets ship with the reference).

The reference's checkpoints (``net_G.pth``, ``vqgan.pkl``) are plain ``state_dict``s
(code/models/base_model.py:93-108).  ``data/state_shapes.json`` lists every key and shape of
``VQLLFLOWDeformable`` (824 tensors), ``VQModel`` (257) and the stage-2 ``LLFlowVQGAN2`` (630) as
dumped from the reference constructors, so a checkpoint with the exact key set can be generated on a
machine where ``/root/reference`` is not mounted.  Zero-initialised reference parameters
(``Conv2dZeros`` flow.py:64-66, ``conv_offset`` ops/dcn/deform_conv.py:367-371, ActNorm
FlowActNorms.py:23-24) are given non-trivial values so every code path is exercised.

Everything is generated with a CPU ``torch.Generator`` in key order -> identical bits on every host
running the same torch build; ``state_fingerprint`` lets a test assert that.
"""
import json
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHAPES = None


def state_shapes(which):
    global _SHAPES
    if _SHAPES is None:
        with open(os.path.join(_HERE, "data", "state_shapes.json")) as f:
            _SHAPES = json.load(f)
    return _SHAPES[which]


def _randn(g, shape):
    return torch.randn(shape, generator=g, dtype=torch.float32)


def _one(key, shape, g):
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith("invconv.weight"):
        q, _ = torch.linalg.qr(_randn(g, shape).double())
        return (q.float() * (1.0 + 0.1 * _randn(g, (shape[0], 1)))).contiguous()
    if key.endswith("actnorm.bias"):
        return 0.1 * _randn(g, shape)
    if key.endswith("actnorm.logs"):
        # outer (3-channel) ActNorms get a positive mean so the 28-step inverse chain stays O(1)
        return 0.1 * _randn(g, shape) + (0.12 if shape[1] == 3 else 0.0)
    if leaf == "logs":                                   # Conv2dZeros.logs [C,1,1]
        return 0.1 * _randn(g, shape)
    if key.endswith("embedding.weight"):
        return _randn(g, shape)
    if ".mix." in key:
        return _randn(g, shape)
    if len(shape) == 1:
        if "norm" in key and leaf == "weight":           # GroupNorm gamma
            return 1.0 + 0.1 * _randn(g, shape)
        if "norm" in key and leaf == "bias":
            return 0.1 * _randn(g, shape)
        if key.endswith("dcn.bias"):
            return 0.5 + 0.1 * _randn(g, shape)
        if key.endswith("conv_offset.bias"):
            return 0.1 * _randn(g, shape)
        return 0.02 * _randn(g, shape)                   # conv biases
    if len(shape) == 4:
        fan_in = shape[1] * shape[2] * shape[3]
        if key.endswith("conv_offset.weight"):
            return (1.5 / math.sqrt(fan_in)) * _randn(g, shape)
        if key.endswith(".4.weight") and ".affine." in key:   # Conv2dZeros of the coupling nets
            return (0.5 / math.sqrt(fan_in)) * _randn(g, shape)
        return (1.0 / math.sqrt(fan_in)) * _randn(g, shape)
    return 0.05 * _randn(g, shape)


def synth_state_dict(which="netG", seed=0):
    """which in {'netG', 'vqgan', 'netG_stage2'} -> OrderedDict[str, fp32 CPU tensor]."""
    g = torch.Generator(device="cpu")
    g.manual_seed({"netG": 1000, "vqgan": 2000, "netG_stage2": 1000}[which] + seed)
    out = {}
    for key, shape in state_shapes(which).items():
        out[key] = _one(key, tuple(shape), g).contiguous()
    return out


def state_fingerprint(sd):
    """Order-dependent fp64 checksum of a state dict (compared with the value stored in the goldens)."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        acc += float(v.double().sum()) * (1 + (i % 7)) + float(v.double().abs().sum())
    return acc


def synth_images(batch, height=400, width=600, seed=0):
    """smooth clean image ``gt`` in [0,1] and its uint8-quantised low-light version
    ``lq`` (gamma 2.2, x0.1, Gaussian read noise).  Returns (lq, gt) float32 [B,3,H,W] in [0,1]."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7000 + seed)
    lo = torch.rand((batch, 3, (height + 7) // 8 + 1, (width + 7) // 8 + 1), generator=g)
    gt = torch.nn.functional.interpolate(lo, scale_factor=8, mode="bilinear", align_corners=False)
    gt = gt[:, :, :height, :width].clamp(0, 1).contiguous()
    noise = torch.randn(gt.shape, generator=g) * (2.0 / 255.0)
    lq = torch.round(255.0 * (0.1 * gt ** 2.2 + noise).clamp(0, 1)) / 255.0
    return lq.contiguous(), gt


def pad_lol(img):
    """infer_dataset_lol.py:124: reflect-pad 20 px at the bottom and the left (400x600 -> 420x620)."""
    return torch.nn.functional.pad(img, (20, 0, 0, 20), mode="reflect")


def _symmetric_index(n, before, after):
    # cv2.BORDER_REFLECT: fedcba|abcdefgh|hgfedcb (edge pixel repeated)
    idx = torch.arange(-before, n + after)
    idx = torch.where(idx < 0, -idx - 1, idx)
    return torch.where(idx >= n, 2 * n - 1 - idx, idx)


def auto_padding(img, times=16):
    """infer_unpaired.py:81-88: pad H and W up to the next multiple of ``times`` (a full ``times``
    when already divisible), split top/bottom and left/right, cv2.BORDER_REFLECT.
    Returns (padded, [h1, h2, w1, w2])."""
    h, w = img.shape[-2:]
    h1, w1 = (times - h % times) // 2, (times - w % times) // 2
    h2, w2 = (times - h % times) - h1, (times - w % times) - w1
    out = img.index_select(-2, _symmetric_index(h, h1, h2)).index_select(-1, _symmetric_index(w, w1, w2))
    return out.contiguous(), [h1, h2, w1, w2]


def preprocess(img01):
    """infer_unpaired.py:121-122 / infer_dataset_lol.py:127-128: log(clamp(x + 1e-3, min=1e-3))"""
    return torch.log(torch.clamp(img01 + 1e-3, min=1e-3))
