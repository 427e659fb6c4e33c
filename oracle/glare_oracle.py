"""CPU fp32 restatement of the GLARE inference / stage-2 hot path.  TEST INFRASTRUCTURE ONLY.

Functional, state-dict driven (no nn.Module tree): every function takes the flat ``state_dict`` of the
reference network it restates plus a key prefix, so the same seeded synthetic checkpoint feeds the
reference (authoring container), this oracle and the CUDA product.  torch-CPU is the fp32 numeric
library for the dense ops (conv2d / group_norm / bmm / softmax -- what the reference itself calls on
CPU); the two operators with exact-index semantics (VQ argmin, DCN sampling) go through the plain-C
restatement in ``oracle_kernels.c``.

Each function cites the reference lines it follows (paths relative to /root/reference/code/models/modules).
Parity pin: ``oracle/gen_golden.py`` compares every stage below with the reference's own modules.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import dcn_im2col, vq_lookup

N_FLOW_STEPS = 28
# FlowUpsamplerNet.py:95-106 with LOL.yml flow.K=12, L=2, additionalFlowNoAffine=2
NO_COUPLING_STEPS = (0, 1, 14, 15)


# ----------------------------------------------------------------------------- taming-style blocks
def _gn(sd, p, x):
    """Normalize = GroupNorm(32, C, eps=1e-6, affine)   encoder_decoder.py:34-35"""
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _swish(x):
    """encoder_decoder.py:29-31"""
    return x * torch.sigmoid(x)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def resnet_block(sd, p, x):
    """ResnetBlock.forward (temb=None, dropout=0)   encoder_decoder.py:117-137"""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


ATTN_MAX_BYTES = 6 << 30      # score matrix budget of attn_block before it switches to blocks of query rows


def attn_block(sd, p, x):
    """AttnBlock.forward: single head, d = C, full softmax over h*w keys   encoder_decoder.py:168-192"""
    hn = _gn(sd, p + ".norm", x)
    q = _conv(sd, p + ".q", hn, padding=0)
    k = _conv(sd, p + ".k", hn, padding=0)
    v = _conv(sd, p + ".v", hn, padding=0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    v = v.reshape(b, c, h * w)
    n = h * w
    if n * n * 4 <= ATTN_MAX_BYTES:
        s = torch.bmm(q, k) * (int(c) ** (-0.5))
        s = torch.softmax(s, dim=2)
        o = torch.bmm(v, s.permute(0, 2, 1)).reshape(b, c, h, w)
    else:
        # 1920x1080 (131 648 tokens): the reference's N x N matrix is 69 GB per sample.  The same arithmetic in blocks of query rows -- each
        # row still sees every key, softmax over the full row -- so only the matmul blocking differs from the one-shot form.
        o = torch.empty((b, c, n), dtype=q.dtype)
        rows = max(1, ATTN_MAX_BYTES // (n * 4))
        for r0 in range(0, n, rows):
            s = torch.bmm(q[:, r0:r0 + rows], k) * (int(c) ** (-0.5))
            s = torch.softmax(s, dim=2)
            o[:, :, r0:r0 + rows] = torch.bmm(v, s.permute(0, 2, 1))
        o = o.reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", o, padding=0)


def downsample(sd, p, x):
    """Downsample: zero-pad right/bottom by 1, 3x3 stride-2 conv   encoder_decoder.py:68-72"""
    return _conv(sd, p + ".conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)


def upsample(sd, p, x):
    """Upsample: nearest x2 then 3x3 conv   encoder_decoder.py:49-53"""
    return _conv(sd, p + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))


def encoder(sd, p, x):
    """Encoder.forward(mid_feat=True), ch_mult (1,2,4), 2 res blocks/level, attention at the lowest
    level   encoder_decoder.py:406-442.  Returns (h, [feat@H (128ch), feat@H/2 (256ch)])."""
    h = _conv(sd, p + ".conv_in", x)
    mid = []
    for lvl in range(3):
        for blk in range(2):
            h = resnet_block(sd, f"{p}.down.{lvl}.block.{blk}", h)
            if f"{p}.down.{lvl}.attn.{blk}.q.weight" in sd:
                h = attn_block(sd, f"{p}.down.{lvl}.attn.{blk}", h)
        if lvl != 2:
            mid.append(h)
            h = downsample(sd, f"{p}.down.{lvl}.downsample", h)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    h = _conv(sd, p + ".conv_out", _swish(_gn(sd, p + ".norm_out", h)))
    return h, mid


def cond_encoder(sd, x, p="RRDB"):
    """ConEncoder1.forward   ConditionEncoder.py:46-55"""
    enc, mid = encoder(sd, p + ".encoder", x)
    return {"cond_feat": torch.sigmoid(_conv(sd, p + ".cond_conv.0", enc)),
            "color_map": _conv(sd, p + ".color_conv", enc),
            "mid_feat": mid, "enc_feat": enc}


def _decoder_trunk(sd, p, z):
    h = _conv(sd, p + ".conv_in", z)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    return h


def vq_decoder_features(sd, zq, p="decoder", pq="post_quant_conv"):
    """VQModel.decode after the quantizer: post_quant_conv 1x1 -> Decoder.forward; only the two
    intermediate feature maps (256ch @H/2, 128ch @H) are consumed downstream
    VQModel_arch.py:88-90, encoder_decoder.py:515-551 (the RGB head :546-550 is discarded by
    VQLLFLOWDeformable_arch.py:246)."""
    h = _decoder_trunk(sd, p, _conv(sd, pq, zq, padding=0))
    feats = []
    for lvl in (2, 1, 0):
        for blk in range(3):
            h = resnet_block(sd, f"{p}.up.{lvl}.block.{blk}", h)
            if lvl == 2:
                h = attn_block(sd, f"{p}.up.{lvl}.attn.{blk}", h)
        if lvl != 2:
            feats.append(h)
        if lvl != 0:
            h = upsample(sd, f"{p}.up.{lvl}.upsample", h)
    return feats


# ----------------------------------------------------------------------------- DCN / AFT
def modulated_deform_conv(x, offset, mask, weight, bias, stride=1, padding=1, dilation=1, dg=4):
    """ModulatedDeformConvFunction.forward: im2col (C oracle) + per-sample GEMM + bias
    ops/dcn/deform_conv.py:124-153, src/deform_conv_cuda.cpp:490-569."""
    B, C, H, W = x.shape
    Co, _, kh, kw = weight.shape
    cols = torch.from_numpy(dcn_im2col(x.numpy(), offset.numpy(), mask.numpy(), kh, kw, stride, padding,
                                       dilation, dg))
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    y = torch.matmul(weight.reshape(Co, -1), cols).reshape(B, Co, Ho, Wo)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    return y


def dcn_offsets_masks(sd, p, feat):
    """DCNv2Pack.forward up to the op call: conv_offset -> chunk3 -> (cat(o1,o2), sigmoid(m))
    deformableDecoder_arch.py:141-149"""
    out = _conv(sd, p + ".conv_offset", feat)
    o1, o2, m = torch.chunk(out, 3, dim=1)
    return torch.cat((o1, o2), dim=1).contiguous(), torch.sigmoid(m).contiguous()


def warp_block(sd, p, x_vq, h, dcn=None):
    """WarpBlock.forward   deformableDecoder_arch.py:285-290.  ``dcn``: a differentiable implementation of the operator (same signature as
    modulated_deform_conv) for gradient checks -- the C im2col restatement has no autograd"""
    feat = _conv(sd, p + ".offset", torch.cat([x_vq, h], dim=1))
    offset, mask = dcn_offsets_masks(sd, p + ".dcn", feat)
    return (dcn or modulated_deform_conv)(x_vq.contiguous(), offset, mask, sd[p + ".dcn.weight"], sd[p + ".dcn.bias"])


def aft_decoder(sd, z, vq_feats, enc_feats, p="deformable_decoder", per_sample_ratio=True, dcn=None):
    """MultiScaleDecoder2.forward   deformableDecoder_arch.py:525-576.
    ``per_sample_ratio``: the reference's ``h.mean()/x_vq.mean()`` (:567) reduces over the whole
    batch but is only ever run at batch 1; per-sample means reproduce that behaviour for any batch
    / sharding (SURVEY.md section 7 'batch coupling')."""
    h = _decoder_trunk(sd, p, z)
    for lvl in (2, 1, 0):
        for blk in range(3):
            h = resnet_block(sd, f"{p}.up.{lvl}.block.{blk}", h)
            if lvl == 2:
                h = attn_block(sd, f"{p}.up.{lvl}.attn.{blk}", h)
        if lvl != 2:
            mixf = torch.sigmoid(sd[f"{p}.mix.{1 - lvl}.w"])
            h = enc_feats[lvl] * mixf + h * (1 - mixf)          # Mix.forward :587-590
            x_vq = warp_block(sd, f"{p}.warp.{1 - lvl}", vq_feats[1 - lvl], h, dcn=dcn)
            if per_sample_ratio:
                ratio = h.mean(dim=(1, 2, 3), keepdim=True) / x_vq.mean(dim=(1, 2, 3), keepdim=True)
            else:
                ratio = h.mean() / x_vq.mean()
            h = h + x_vq * ratio
        if lvl != 0:
            h = upsample(sd, f"{p}.up.{lvl}.upsample", h)
    return _conv(sd, p + ".residual_conv", _swish(_gn(sd, p + ".norm_out", h)))


# ----------------------------------------------------------------------------- flow
def _flow_nn(sd, p, x):
    """CondAffineSeparatedAndCond.F: Conv2d 3x3 (no bias) + ActNorm -> ReLU -> Conv2d 1x1 + ActNorm ->
    ReLU -> Conv2dZeros 3x3 * exp(3*logs)   FlowAffineCouplingsAblation.py:143-151, flow.py:48-70,
    FlowActNorms.py:48-72 (forward: (x + bias) * exp(logs))"""
    h = F.conv2d(x, sd[p + ".0.weight"], None, padding=1)
    h = torch.relu((h + sd[p + ".0.actnorm.bias"]) * torch.exp(sd[p + ".0.actnorm.logs"]))
    h = F.conv2d(h, sd[p + ".2.weight"], None, padding=0)
    h = torch.relu((h + sd[p + ".2.actnorm.bias"]) * torch.exp(sd[p + ".2.actnorm.logs"]))
    h = F.conv2d(h, sd[p + ".4.weight"], sd[p + ".4.bias"], padding=1)
    return h * torch.exp(sd[p + ".4.logs"] * 3)


def _scale_shift(h):
    """thops.split_feature('cross') + sigmoid(scale+2)+eps   FlowAffineCouplingsAblation.py:121-135"""
    return torch.sigmoid(h[:, 1::2] + 2.0) + 0.0001, h[:, 0::2]


def invconv_inverse_weight(w):
    """Permutations.py:38: inverse in fp64, cast to fp32"""
    return torch.inverse(w.double()).float()


def flow_step_inverse(sd, p, z, ft, coupling):
    """FlowStep.reverse_flow   FlowStep.py:100-119"""
    if coupling:
        z1, z2 = z[:, :1], z[:, 1:]
        scale, shift = _scale_shift(_flow_nn(sd, p + ".affine.fAffine", torch.cat([z1, ft], dim=1)))
        z2 = z2 / scale - shift                              # FlowAffineCouplingsAblation.py:88-91
        z = torch.cat([z1, z2], dim=1)
        scale_ft, shift_ft = _scale_shift(_flow_nn(sd, p + ".affine.fFeatures", ft))
        z = z / scale_ft - shift_ft                          # :106-108
    winv = invconv_inverse_weight(sd[p + ".invconv.weight"])
    z = F.conv2d(z, winv.view(3, 3, 1, 1))                   # Permutations.py:55-56
    z = z * torch.exp(-sd[p + ".actnorm.logs"]) - sd[p + ".actnorm.bias"]   # FlowActNorms.py:64,98-99
    return z


def flow_step_forward(sd, p, z, ft, logdet, coupling):
    """FlowStep.normal_flow   FlowStep.py:75-98.  logdet is per-sample [B]."""
    pixels = z.shape[2] * z.shape[3]
    z = (z + sd[p + ".actnorm.bias"]) * torch.exp(sd[p + ".actnorm.logs"])
    logdet = logdet + sd[p + ".actnorm.logs"].sum() * pixels                 # FlowActNorms.py:66-74
    w = sd[p + ".invconv.weight"]
    z = F.conv2d(z, w.view(3, 3, 1, 1))
    logdet = logdet + torch.slogdet(w)[1] * pixels                           # Permutations.py:27,51-53
    if coupling:
        scale_ft, shift_ft = _scale_shift(_flow_nn(sd, p + ".affine.fFeatures", ft))
        z = (z + shift_ft) * scale_ft                                        # :56-59
        logdet = logdet + torch.log(scale_ft).sum(dim=(1, 2, 3))
        z1, z2 = z[:, :1], z[:, 1:]
        scale, shift = _scale_shift(_flow_nn(sd, p + ".affine.fAffine", torch.cat([z1, ft], dim=1)))
        z2 = (z2 + shift) * scale                                            # :75-78
        logdet = logdet + torch.log(scale).sum(dim=(1, 2, 3))
        z = torch.cat([z1, z2], dim=1)
    return z, logdet


def flow_decode(sd, z, ft, p="flowUpsamplerNet", trace=None):
    """FlowUpsamplerNet.decode: steps 27 -> 0   FlowUpsamplerNet.py:290-326"""
    for s in range(N_FLOW_STEPS - 1, -1, -1):
        z = flow_step_inverse(sd, f"{p}.layers.{s}", z, ft, s not in NO_COUPLING_STEPS)
        if trace is not None:
            trace.append(z)
    return z


def flow_encode(sd, gt, ft, logdet=None, p="flowUpsamplerNet"):
    """FlowUpsamplerNet.encode: steps 0 -> 27   FlowUpsamplerNet.py:228-274"""
    z = gt
    if logdet is None:
        logdet = torch.zeros(gt.shape[0])
    for s in range(N_FLOW_STEPS):
        z, logdet = flow_step_forward(sd, f"{p}.layers.{s}", z, ft, logdet, s not in NO_COUPLING_STEPS)
    return z, logdet


def stage2_nll(sd, gt_latent, lr, quant=32, p_enc="RRDB", use_gt_mean=False):
    """LLFlowVQGAN2.normal_flow   LLFlowVQGAN2_arch.py:75-122 (add_gt_noise=False path as called by LLFlow_model.py:215).
    use_gt_mean: the branch `random.random() > train_gt_ratio` picks at :109 -- False: mean = color_map, True: mean = gt.
    Returns (z, nll[B])."""
    enc = cond_encoder(sd, lr, p_enc)
    pixels = gt_latent.shape[2] * gt_latent.shape[3]
    z, logdet = flow_encode(sd, gt_latent, enc["cond_feat"])
    mean = gt_latent if use_gt_mean else enc["color_map"]
    logp = (-0.5 * ((z - mean) ** 2 + math.log(2 * math.pi))).sum(dim=(1, 2, 3))   # flow.py:76-95
    nll = -(logdet + logp) / float(np.log(2.) * pixels)
    return z, nll


# ----------------------------------------------------------------------------- VQ + full pipeline
def vector_quantize(sd_vq, z):
    """VectorQuantizer2.forward (inference outputs)   quantize.py:271-312"""
    idx, zq = vq_lookup(z.numpy(), sd_vq["quantize.embedding.weight"].numpy())
    return torch.from_numpy(zq), torch.from_numpy(idx)


def glare_infer(sd_g, sd_vq, lr, per_sample_ratio=True, stages=None):
    """VQLLFLOWDeformable.reverse_flow   VQLLFLOWDeformable_arch.py:222-250.
    lr: [B,3,H,W] = log(clamp(x+1e-3)) pre-processed low-light input.  Returns the RGB prediction
    (un-clamped) and, when ``stages`` is a dict, every intermediate used for teacher-forced parity."""
    with torch.no_grad():
        enc = cond_encoder(sd_g, lr)
        z = flow_decode(sd_g, enc["color_map"], enc["cond_feat"])
        zq, idx = vector_quantize(sd_vq, z)
        vq_feats = vq_decoder_features(sd_vq, zq)
        out = aft_decoder(sd_g, z, vq_feats, enc["mid_feat"], per_sample_ratio=per_sample_ratio)
    if stages is not None:
        stages.update(cond_feat=enc["cond_feat"], color_map=enc["color_map"], mid0=enc["mid_feat"][0],
                      mid1=enc["mid_feat"][1], z_flow=z, z_q=zq, idx=idx, vq_feat1=vq_feats[0],
                      vq_feat0=vq_feats[1], out=out)
    return out


def preprocess(img01):
    """infer_unpaired.py:121-122 / infer_dataset_lol.py:127-128: log(clamp(x + 1e-3, min=1e-3))"""
    return torch.log(torch.clamp(img01 + 1e-3, min=1e-3))


def psnr(a, b):
    """utils/utils2.py:32-36 on [0,1] images"""
    mse = torch.mean((a - b) ** 2)
    return float(10 * torch.log10(1.0 / mse))
