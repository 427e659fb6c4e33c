"""CPU restatement of the stage-3 losses (TEST INFRASTRUCTURE: only tests/, smoke and bench's cpu_baseline may import this package).

    msssim        code/models/modules/pytorch_msssim/__init__.py:7-97
    perceptual    code/models/modules/losses.py:12-40 over torchvision vgg16.features[:16]
    stage3_loss   code/models/VQLLFLOWD_model.py:212-223

Plain torch, differentiable by autograd (the gradient reference for glare_b200/losses.py).  Pinned against the reference's own functions by
tests/test_losses_cpu.py::test_oracle_losses_match_the_reference and oracle/gen_golden_stage3.py (tests/golden/stage3.npz)."""
import math

import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def window(ws, channels, sigma=1.5):
    g = torch.tensor([math.exp(-(i - ws // 2) ** 2 / float(2 * sigma ** 2)) for i in range(ws)], dtype=torch.float32)     # :7-9
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float()[None, None].expand(channels, 1, ws, ws).contiguous()                                  # :12-17


def ssim_level(x, y, window_size=11, val_range=None):
    """-> (mean ssim_map, mean cs_map)   :20-68 with size_average=True, full=True"""
    if val_range is None:
        L = (255 if float(x.detach().max()) > 128 else 1) - (-1 if float(x.detach().min()) < -0.5 else 0)
    else:
        L = val_range
    C, H, W = x.shape[1:]
    w = window(min(window_size, H, W), C).to(x.dtype)
    blur = lambda t: F.conv2d(t, w, padding=0, groups=C)                  # noqa: E731
    mu1, mu2 = blur(x), blur(y)
    s11, s22, s12 = blur(x * x) - mu1 * mu1, blur(y * y) - mu2 * mu2, blur(x * y) - mu1 * mu2
    C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    v1, v2 = 2.0 * s12 + C2, s11 + s22 + C2
    ssim_map = ((2 * mu1 * mu2 + C1) * v1) / ((mu1 * mu1 + mu2 * mu2 + C1) * v2)
    return ssim_map.mean(), (v1 / v2).mean()


def msssim(x, y, window_size=11, val_range=None, normalize=False):
    """:71-97"""
    wts = torch.tensor(WEIGHTS, dtype=x.dtype)
    sims, css = [], []
    for _ in WEIGHTS:
        s, c = ssim_level(x, y, window_size, val_range)
        sims.append(s)
        css.append(c)
        x, y = F.avg_pool2d(x, (2, 2)), F.avg_pool2d(y, (2, 2))
    sims, css = torch.stack(sims), torch.stack(css)
    if normalize:
        sims, css = (sims + 1) / 2, (css + 1) / 2
    return torch.prod((css ** wts)[:-1] * (sims ** wts)[-1])              # :96 as written (not the textbook product)


VGG_LAYERS = ("c0", "r", "c2", "r*", "p", "c5", "r", "c7", "r*", "p", "c10", "r", "c12", "r", "c14", "r*")      # torchvision vgg16.features[:16]


def vgg_features(sd, x):
    """relu1_2, relu2_2, relu3_3 (losses.py:20-33); sd keys '<idx>.weight' / '<idx>.bias'"""
    out = []
    for op in VGG_LAYERS:
        if op[0] == "c":
            x = F.conv2d(x, sd[op[1:] + ".weight"], sd[op[1:] + ".bias"], padding=1)
        elif op[0] == "p":
            x = F.max_pool2d(x, 2, 2)
        else:
            x = F.relu(x)
            if op.endswith("*"):
                out.append(x)
    return out


def perceptual(sd, x, gt):
    """losses.py:35-40"""
    terms = [F.mse_loss(a, b) for a, b in zip(vgg_features(sd, x), vgg_features(sd, gt))]
    return sum(terms) / len(terms)


def stage3_loss(rec, real_H, vgg_sd):
    """VQLLFLOWD_model.py:212-223 -> (total, dict of the three terms)"""
    sr = rec.to(torch.float32).clamp(0, 1)
    ok = ~torch.isnan(sr)
    sr = torch.where(ok, sr, torch.zeros_like(sr))
    terms = {"l1_loss": ((sr - real_H) * ok).abs().mean(), "percep_loss": perceptual(vgg_sd, sr, real_H) * 0.01,
             "ssim_loss": (1 - msssim(sr, real_H, normalize=True)) * 0.2}
    return sum(terms.values()), terms
