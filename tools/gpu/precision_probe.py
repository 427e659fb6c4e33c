"""Teacher-forced precision probe of the post-VQ decoders (per-stage precision plan).

The VQ index is decided before the decoders, so the two decoders (VQGAN decoder + AFT/DCN decoder, ~80 % of the conv FLOPs, 7 of the 11
attention blocks) only have to hold the 1e-3 pixel bar given the SAME z / z_q / encoder features.  This probe feeds one 420x620 image's
pre-VQ stage outputs (default fp32-grade backend) into the decoders under cheaper operand treatments and reports the pixel error against
the cuDNN fp32 (TF32 off) library path:

  mode4            bf16x3 everywhere (today's default)
  mode1            tf32 single pass (operands truncated to 11 bits by the tensor core)
  mode0            bf16 single pass
  w-fp16 + mode4   decoder conv weights rounded to fp16 (11-bit significand) and run in bf16x3: the split of such a weight is exact, so this
                   is the arithmetic of a TWO-pass scheme "activation split in two pieces x single fp16 weight piece"
  w-bf16 + mode4   the same with bf16 weights (two-pass scheme with bf16 pieces)
"""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from glare_b200 import synth  # noqa: E402
from glare_b200.dense import make_dense  # noqa: E402
from glare_b200.engine import GlareEngine  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 600)
dev = "cuda:0"
sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
lq, gt = synth.synth_images(1, H, W, seed=0)
lr = synth.preprocess(synth.pad_lol(lq))


def rounded(sd, prefixes, fn):
    return {k: (fn(v) if v.dim() == 4 and k.startswith(prefixes) else v) for k, v in sd.items()}


def fp16(v):
    return v.half().float()


def bf16(v):
    return v.bfloat16().float()


with torch.no_grad():
    e4 = GlareEngine(sd_g, sd_v, device=dev, dense=make_dense("auto"))
    st = {}
    e4.infer(lr, stages=st)
    z, zq, mids = st["z_flow"], st["z_q"], [st["mid0"], st["mid1"]]

    def decoders(eng, which="both", base=None):
        """which: 'both', 'vq' (only the VQGAN decoder in the variant, the AFT decoder from `base`), 'aft'"""
        e_vq = eng if which in ("both", "vq") else base
        e_aft = eng if which in ("both", "aft") else base
        feats = e_vq.vq_decoder_features(zq)
        return e_aft.aft_decoder(z, [f.float() for f in feats], mids).float()

    from libdense import TorchDense
    ref = decoders(GlareEngine(sd_g, sd_v, device=dev, dense=TorchDense()))
    crop = lambda o: o[:, :, :H, 20:].clamp(0, 1)           # noqa: E731  (pad_lol: 20 px left / bottom)

    def report(name, out):
        d = (crop(out) - crop(ref)).abs()
        print("%-44s pixel max %.3e  mean %.3e  p99.99 %.3e  (unclamped max %.3e)" %
              (name, float(d.max()), float(d.mean()), float(d.flatten().kthvalue(int(d.numel() * 0.9999)).values),
               float((out - ref).abs().max())), flush=True)

    report("mode4 bf16x3", decoders(e4))
    pre_v, pre_g = ("decoder.", "post_quant_conv."), ("deformable_decoder.",)
    variants = [("mode1 tf32", "tc-tf32", None), ("mode3 tf32+2xbf16", "tc-tf32bf16x2", None), ("mode0 bf16", "tc-bf16", None),
                ("w-fp16 + mode4 (2-pass fp16 scheme)", "auto", fp16), ("w-bf16 + mode4 (2-pass bf16 scheme)", "auto", bf16)]
    for name, dn, fn in variants:
        g2 = rounded(sd_g, pre_g, fn) if fn else sd_g
        v2 = rounded(sd_v, pre_v, fn) if fn else sd_v
        eng = GlareEngine(g2, v2, device=dev, dense=make_dense(dn))
        report(name, decoders(eng))
        if dn != "tc-bf16":
            report(name + " [VQ decoder only]", decoders(eng, "vq", e4))
            report(name + " [AFT decoder only]", decoders(eng, "aft", e4))
        del eng
        torch.cuda.empty_cache()
