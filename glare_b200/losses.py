"""Stage-3 losses on the library's kernels, with the reference's names and call signatures (VQLLFLOWD_model.py:89-90, 216-223):

    PerceptualNetwork        code/models/modules/losses.py:12-40   VGG16 features[:16] (relu1_2, relu2_2, relu3_3), mean of the three MSEs
    msssim(img1, img2, ...)  code/models/modules/pytorch_msssim/__init__.py:71-97 (per level: `ssim` :20-68)

Both are single autograd nodes with a gradient for the FIRST image only (the reconstruction; the second is the ground truth), so the
reference's ``total_loss.backward()`` runs unchanged on top of them.  The VGG convolutions run on the tensor-core conv path through the same
tape as the encoder / decoder training code (its weights are frozen: no weight gradients); the SSIM levels, ReLU, max-pool and the
avg_pool2d pyramid are csrc/loss.cu.

The reference constructs ``vgg16(pretrained=True)`` (a download).  PerceptualNetwork takes the weights as a state dict with torchvision's
keys (``features.N.weight`` or ``N.weight``), from the file named by GLARE_VGG16_WEIGHTS, or -- ``pretrained=True``, what the drop-in binding
uses -- from torchvision like the reference; otherwise it initialises like torchvision's ``vgg16(weights=None)`` (tests and the bench: there
is no network on the build / GPU boxes).
"""
import ctypes
import math

import torch
import torch.nn as nn

from .encoder_train import BlockGraph

VGG_CONVS = ((0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256))   # torchvision vgg16.features[:16]
VGG_TAPS = (3, 8, 15)                                                                                                # losses.py:20-24
VGG_POOLS = (4, 9)


# ------------------------------------------------------------------------------------------------------------ MS-SSIM
def gaussian_window(ws, sigma=1.5):
    """pytorch_msssim/__init__.py:7-9, evaluated like the reference (python floats -> fp32 tensor -> normalised in fp32)"""
    g = torch.tensor([math.exp(-(x - ws // 2) ** 2 / float(2 * sigma ** 2)) for x in range(ws)], dtype=torch.float32)
    return g / g.sum()


class SsimKernels:
    """one SSIM level on the GPU (csrc/loss.cu)"""

    def __init__(self):
        from . import ops
        from ._lib import lib, stream
        self.ops, self.lib, self.stream = ops, lib, stream

    def _call(self, name, *args):
        self.ops.check(getattr(self.lib(), name)(*args, self.stream()), name)

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _win(self, win):
        return (ctypes.c_float * len(win))(*[float(v) for v in win])

    def _partials(self, planes, H, W, ws):
        fn = self.lib().glare_ssim_partials
        return int(fn(planes, H, W, ws))

    def fwd(self, x, y, win, C1, C2):
        """x, y [B,C,H,W] fp32 contiguous -> (sum of cs_map, sum of ssim_map) as fp64 device scalars"""
        B, C, H, W = x.shape
        ws = len(win)
        part = torch.empty((self._partials(B * C, H, W, ws), 2), device=x.device, dtype=torch.float32)
        self._call("glare_ssim_fwd_f32", self._p(x), self._p(y), B * C, H, W, ws, self._win(win), ctypes.c_float(C1), ctypes.c_float(C2),
                   self._p(part))
        s = part.double().sum(dim=0)
        return s[0], s[1]

    def avgpool2(self, x):
        B, C, H, W = x.shape
        y = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=torch.float32)
        self._call("glare_avgpool2_f32", self._p(x), B * C, H, W, self._p(y))
        return y

    def bwd(self, x, y, win, C1, C2, coef, coarse):
        B, C, H, W = x.shape
        ws = len(win)
        maps = torch.empty((3, B * C, H - ws + 1, W - ws + 1), device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        self._call("glare_ssim_bwd_f32", self._p(x), self._p(y), B * C, H, W, ws, self._win(win), ctypes.c_float(C1), ctypes.c_float(C2),
                   self._p(coef), self._p(maps[0]), self._p(maps[1]), self._p(maps[2]), self._p(coarse), self._p(dx))
        return dx


MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)             # pytorch_msssim/__init__.py:73


def _combine(mssim, mcs, normalize):
    """pytorch_msssim/__init__.py:86-97, including its `prod(pow1[:-1] * pow2[-1])` (the last level's term enters once per coarser level)"""
    weights = torch.tensor(MSSSIM_WEIGHTS, device=mssim.device, dtype=mssim.dtype)
    if normalize:
        mssim, mcs = (mssim + 1) / 2, (mcs + 1) / 2
    pow1, pow2 = mcs ** weights, mssim ** weights
    return torch.prod(pow1[:-1] * pow2[-1])


def _value_range(img):
    """`ssim` with val_range=None (:22-33): decided from img1 per level, on the host like the reference's `if torch.max(img1) > 128`"""
    mx, mn = float(img.max()), float(img.min())
    return (255 if mx > 128 else 1) - (-1 if mn < -0.5 else 0)


class MsssimFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, img1, img2, kernels, window_size, val_range, normalize):
        x, y = img1.detach().float().contiguous(), img2.detach().float().contiguous()
        levels = len(MSSSIM_WEIGHTS)
        tape, sims, css = [], [], []
        for lvl in range(levels):
            B, C, H, W = x.shape
            ws = min(window_size, H, W)                                    # real_size :37
            L = _value_range(x) if val_range is None else val_range
            C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
            win = gaussian_window(ws).tolist()
            n_pos = float(B * C * (H - ws + 1) * (W - ws + 1))
            s_cs, s_ss = kernels.fwd(x, y, win, C1, C2)
            css.append(s_cs / n_pos)
            sims.append(s_ss / n_pos)
            tape.append((x, y, win, C1, C2, n_pos))
            if lvl + 1 < levels:
                x, y = kernels.avgpool2(x), kernels.avgpool2(y)            # :83-84
        with torch.enable_grad():
            cs_t = torch.stack(css).float().requires_grad_(True)
            ss_t = torch.stack(sims).float().requires_grad_(True)
            out = _combine(ss_t, cs_t, normalize)
            g_cs, g_ss = torch.autograd.grad(out, (cs_t, ss_t))
        ctx.tape, ctx.kernels = tape, kernels
        ctx.coefs = torch.stack([g_cs, g_ss], dim=1)                      # [levels][2]
        return out.detach().to(img1.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        if ctx.tape is None:
            raise RuntimeError("Trying to backward through msssim a second time: its saved levels are freed by the first backward pass")
        dx = None
        for lvl in reversed(range(len(ctx.tape))):
            x, y, win, C1, C2, n_pos = ctx.tape[lvl]
            coef = (ctx.coefs[lvl] * (g_out.float() / n_pos)).contiguous()
            dx = ctx.kernels.bwd(x, y, win, C1, C2, coef, dx)
        ctx.tape = None
        return dx, None, None, None, None, None


_KERNELS = None


def msssim(img1, img2, window_size=11, size_average=True, val_range=None, normalize=False, kernels=None):
    """pytorch_msssim.msssim (:71-97); differentiable in img1"""
    global _KERNELS
    if not size_average:
        raise NotImplementedError("size_average=False is not used by the reference's stage 3 (VQLLFLOWD_model.py:221)")
    if img2.requires_grad:
        raise NotImplementedError("msssim is differentiable in its first argument only (the second is the ground truth)")
    if kernels is None:
        if not img1.is_cuda:
            raise NotImplementedError("glare_b200.losses.msssim runs on CUDA tensors")
        if _KERNELS is None:
            _KERNELS = SsimKernels()
        kernels = _KERNELS
    return MsssimFn.apply(img1, img2, kernels, window_size, val_range, normalize)


# ------------------------------------------------------------------------------------------------------------ VGG perceptual
class VGGGraph(BlockGraph):
    """vgg16.features[:16] on the tape; parameter keys '<idx>.weight' / '<idx>.bias'"""

    def forward(self, x):
        self._begin(x)
        convs = {i: None for i, _, _ in VGG_CONVS}
        h, taps = 0, []
        self.switches = []                              # ReLU outputs and pooling indices: the discontinuous choices of the forward pass
        for idx in range(16):
            if idx in convs:
                h = self._conv(str(idx), h, need_gx=True, need_gw=False)
            elif idx in VGG_POOLS:
                y, saved = self.L.maxpool2(self.vals[h])
                self.switches.append(saved[0])
                h = self._fn(y, (h,), lambda gy, saved=saved: (self.L.maxpool2_bwd(gy, saved),))
            else:
                y = self.L.relu(self.vals[h])
                self.switches.append(y)
                h = self._fn(y, (h,), lambda gy, y=y: (self.L.relu_bwd(y, gy),))
                if idx in VGG_TAPS:
                    taps.append(h)
        self.taps = taps
        return [self.vals[t] for t in taps]

    def backward(self, g_taps):
        g = self._backprop({t: gt for t, gt in zip(self.taps, g_taps)})
        self.release()
        return g[0]


class PerceptualFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, dehaze, gt, leaves, sd):
        x = dehaze.detach().float()
        with torch.no_grad():
            g_gt = VGGGraph(leaves, sd)
            f_gt = g_gt.forward(gt.detach().float())
            g_gt.release()
            graph = VGGGraph(leaves, sd)
            f_x = graph.forward(x)
            seeds, loss = [], 0.0
            for a, b in zip(f_x, f_gt):
                d = a - b
                loss = loss + (d * d).mean()                               # F.mse_loss :38
                seeds.append(d * (2.0 / (d.numel() * len(f_x))))
        ctx.graph, ctx.seeds = graph, seeds
        return (loss / len(f_x)).to(dehaze.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        if ctx.graph is None:
            raise RuntimeError("Trying to backward through the perceptual loss a second time: its tape is freed by the first backward pass")
        with torch.no_grad():
            gx = ctx.graph.backward([s * g_out.float() for s in ctx.seeds])
        ctx.graph = ctx.seeds = None
        return gx, None, None, None


class PerceptualNetwork(nn.Module):
    """losses.py:12-40.  ``vgg_model`` holds the seven conv layers of torchvision's vgg16.features[:16] under the reference's keys
    (``vgg_model.<idx>.weight``), frozen."""

    def __init__(self, state_dict=None, leaves=None, pretrained=False):
        super().__init__()
        import os
        if state_dict is None and os.environ.get("GLARE_VGG16_WEIGHTS"):
            state_dict = torch.load(os.environ["GLARE_VGG16_WEIGHTS"], map_location="cpu")       # torchvision's vgg16 state dict
        # pretrained=True: torchvision's weights, the reference's own source (losses.py:15) -- fetched at the first loss evaluation, not at
        # construction: the reference builds this module for inference too (VQLLFLOWD_model.py:89), where it is never called
        self._fetch_pretrained = state_dict is None and pretrained
        self.vgg_model = nn.Module()
        for idx, ci, co in VGG_CONVS:
            m = nn.Module()
            w = torch.empty((co, ci, 3, 3))
            nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")          # torchvision vgg16(weights=None)
            m.register_parameter("weight", nn.Parameter(w, requires_grad=False))
            m.register_parameter("bias", nn.Parameter(torch.zeros(co), requires_grad=False))
            self.vgg_model.add_module(str(idx), m)
        self.layer_name_mapping = {"3": "relu1_2", "8": "relu2_2", "15": "relu3_3"}
        self._leaves = leaves
        if state_dict is not None:
            sd = {}
            for k, v in state_dict.items():
                k = k[len("features."):] if k.startswith("features.") else k
                k = k[len("vgg_model."):] if k.startswith("vgg_model.") else k
                if k.split(".")[0].isdigit() and int(k.split(".")[0]) < 16:
                    sd[k] = v
            self.vgg_model.load_state_dict(sd, strict=True)

    def leaves(self, device):
        if self._leaves is None:
            if device.type != "cuda":
                raise NotImplementedError("glare_b200.losses.PerceptualNetwork runs on CUDA tensors")
            from .dense import make_dense
            from .encoder_train import CudaLeaves
            self._leaves = CudaLeaves(make_dense("auto"))
        return self._leaves

    def _sd(self):
        if self._fetch_pretrained:
            from torchvision.models import vgg16
            sd = {k: v for k, v in vgg16(pretrained=True).features.state_dict().items() if int(k.split(".")[0]) < 16}
            self.vgg_model.load_state_dict(sd, strict=True)
            self._fetch_pretrained = False
        # the weights are frozen: hand the SAME tensor objects to every call, so that the conv path's packed-weight cache and the flipped
        # filters of the data gradients (both keyed on the tensor they were made from) are built once, not once per loss evaluation
        key = tuple((p.data_ptr(), p._version) for p in self.vgg_model.parameters())
        if getattr(self, "_sd_key", None) != key:
            self._sd_key, self._sd_cache = key, {k: v.detach() for k, v in self.vgg_model.state_dict().items()}
        return self._sd_cache

    @torch.no_grad()
    def output_features(self, x):
        graph = VGGGraph(self.leaves(x.device), self._sd())
        feats = graph.forward(x.float())
        graph.release()
        return feats

    def forward(self, dehaze, gt):
        return PerceptualFn.apply(dehaze, gt, self.leaves(dehaze.device), self._sd())


def stage3_loss(sr_raw, real_H, perceptual, msssim_fn=msssim):
    """the objective of VQLLFLOWDModel.optimize_parameters (VQLLFLOWD_model.py:212-223) on a reconstruction ``sr_raw`` (``rec`` there):
    clamp, NaN masking, |sr - gt| mean + 0.01 * perceptual + 0.2 * (1 - MS-SSIM).  Returns (total, {name: term})."""
    sr = sr_raw.to(torch.float32).clamp(0, 1)
    not_nan = ~torch.isnan(sr)
    sr = torch.where(not_nan, sr, torch.zeros_like(sr))
    losses = {"l1_loss": ((sr - real_H) * not_nan).abs().mean(),
              "percep_loss": perceptual(sr, real_H) * 0.01,
              # val_range=1: what `ssim` derives from a [0, 1]-clamped img1 at every level (:22-33), without its two host reads per level
              "ssim_loss": (1 - msssim_fn(sr, real_H, normalize=True, val_range=1)) * 0.2}
    return sum(losses.values()), losses
