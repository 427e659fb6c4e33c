"""Pin the oracle against the UNMODIFIED reference and write the golden vectors.  TEST INFRASTRUCTURE.

Run in the authoring container only (needs /root/reference):

    python -m oracle.gen_golden            # writes tests/golden/*.npz and tests/golden/PIN_REPORT.txt

For every stage of the hot path this script (1) runs the reference's own Python module
(imported through oracle/ref_shims.py) on seeded synthetic weights/inputs, (2) runs the oracle
restatement (oracle/glare_oracle.py, oracle/oracle_kernels.c) on the same data, (3) records the
difference in the report and (4) stores the REFERENCE's outputs as goldens.  The CUDA product is
then compared against those goldens (tests/test_*_gpu.py) and the oracle against them on CPU
(tests/test_oracle.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from glare_b200 import synth  # noqa: E402  (host utility: seeded weights; not a compute path)
from oracle import glare_oracle as O  # noqa: E402
from oracle import ref_shims, vq_lookup, dcn_im2col  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REPORT = []


def log(msg):
    print(msg)
    REPORT.append(msg)


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    netG, net_hq, opt = ref_shims.build_reference("LOL.yml", seed=0)
    sd_g = synth.synth_state_dict("netG", 0)
    sd_v = synth.synth_state_dict("vqgan", 0)
    netG.load_state_dict(sd_g, strict=True)
    net_hq.load_state_dict(sd_v, strict=True)
    netG.eval(), net_hq.eval()
    fp_g, fp_v = synth.state_fingerprint(sd_g), synth.state_fingerprint(sd_v)
    log("state fingerprints: netG %.10e  vqgan %.10e" % (fp_g, fp_v))
    g = torch.Generator().manual_seed(1234)

    # ------------------------------------------------------------------ 1. VQ (quantize.py:271-312)
    cb = sd_v["quantize.embedding.weight"]
    z = torch.randn((2, 3, 37, 53), generator=g) * 1.2
    # adversarial tokens: exact codebook entries, midpoints between two codes (near ties), huge values
    z[0, :, 0, :8] = cb[:8].t()
    z[0, :, 1, :8] = ((cb[10:18] + cb[20:28]) * 0.5).t()
    z[1, :, 0, 0] = torch.tensor([50.0, -40.0, 30.0])
    z[1, :, 0, 1] = 0.0
    with torch.no_grad():
        zq_ref, _, (_, _, idx_ref) = net_hq.quantize(z)
    idx_o, zq_o = vq_lookup(z.numpy(), cb.numpy())
    n_bad = int((torch.from_numpy(idx_o) != idx_ref).sum())
    log("VQ small: oracle vs reference index mismatches %d / %d ; zq maxdiff %.3g" %
        (n_bad, idx_ref.numel(), maxdiff(torch.from_numpy(zq_o), zq_ref)))
    assert n_bad == 0
    # a duplicate-code codebook exercises the first-index tie-break
    cb_dup = cb.clone()
    cb_dup[4096:] = cb[:4096]
    net_hq.quantize.embedding.weight.data.copy_(cb_dup)
    with torch.no_grad():
        _, _, (_, _, idx_dup) = net_hq.quantize(z)
    net_hq.quantize.embedding.weight.data.copy_(cb)
    idx_o2, _ = vq_lookup(z.numpy(), cb_dup.numpy())
    assert int((torch.from_numpy(idx_o2) != idx_dup).sum()) == 0 and int(idx_dup.max()) < 4096
    log("VQ duplicate-codebook tie-break: oracle == reference, max idx %d" % int(idx_dup.max()))
    # large: latent-size token grid (LOL eval shape 105x155) -- bit-exactness at scale
    zl = torch.randn((1, 3, 105, 155), generator=g)
    with torch.no_grad():
        _, _, (_, _, idx_l) = net_hq.quantize(zl)
    idx_lo, _ = vq_lookup(zl.numpy(), cb.numpy())
    n_bad = int((torch.from_numpy(idx_lo) != idx_l).sum())
    log("VQ 105x155: oracle vs reference index mismatches %d / %d" % (n_bad, idx_l.numel()))
    assert n_bad == 0
    save("vq", z=z, idx=idx_ref, zq=zq_ref, idx_dup=idx_dup, z_large=zl, idx_large=idx_l.to(torch.int32),
         fingerprint_vqgan=np.float64(fp_v))

    # ------------------------------------------------------------------ 2. flow steps (FlowStep.py:75-119)
    flow = netG.flowUpsamplerNet
    h, w = 12, 20
    zf = torch.randn((2, 3, h, w), generator=g)
    ft = torch.sigmoid(torch.randn((2, 64, h, w), generator=g))
    gold = {"z": zf, "ft": ft}
    for s in (0, 2, 15, 27):
        layer = flow.layers[s]
        coupling = s not in O.NO_COUPLING_STEPS
        with torch.no_grad():
            z_inv_ref, _ = layer(zf, logdet=torch.zeros(2), reverse=True, rrdbResults=ft)
            z_fwd_ref, ld_ref = layer(zf, logdet=torch.zeros(2), reverse=False, rrdbResults=ft)
            z_inv_o = O.flow_step_inverse(sd_g, "flowUpsamplerNet.layers.%d" % s, zf, ft, coupling)
            z_fwd_o, ld_o = O.flow_step_forward(sd_g, "flowUpsamplerNet.layers.%d" % s, zf, ft, torch.zeros(2), coupling)
        log("flow step %2d (%s): inverse maxdiff %.3g  forward maxdiff %.3g  logdet diff %.3g (|ld| %.3g)" %
            (s, "coupling" if coupling else "noCoupling", maxdiff(z_inv_o, z_inv_ref), maxdiff(z_fwd_o, z_fwd_ref),
             maxdiff(ld_o, ld_ref), float(ld_ref.abs().max())))
        assert maxdiff(z_inv_o, z_inv_ref) < 1e-5 and maxdiff(z_fwd_o, z_fwd_ref) < 1e-5
        gold["inv_%d" % s] = z_inv_ref
        gold["fwd_%d" % s] = z_fwd_ref
        gold["logdet_%d" % s] = ld_ref
    # whole chain, both directions (FlowUpsamplerNet.py:228-326)
    with torch.no_grad():
        x_ref, _ = flow(rrdbResults={"cond_feat": ft}, z=zf, reverse=True, logdet=torch.zeros(2), eps_std=None, epses=None)
        x_o = O.flow_decode(sd_g, zf, ft)
        zz_ref, ld_ref = flow(rrdbResults={"cond_feat": ft}, gt=zf, reverse=False, logdet=torch.zeros(2), epses=None)
        zz_o, ld_o = O.flow_encode(sd_g, zf, ft)
    log("flow decode 28 steps: maxdiff %.3g (|x| max %.3g); encode: maxdiff %.3g (|z| max %.3g) logdet diff %.3g (|ld| %.3g)" %
        (maxdiff(x_o, x_ref), float(x_ref.abs().max()), maxdiff(zz_o, zz_ref), float(zz_ref.abs().max()),
         maxdiff(ld_o, ld_ref), float(ld_ref.abs().max())))
    gold.update(decode=x_ref, encode=zz_ref, encode_logdet=ld_ref)
    save("flow", fingerprint_netG=np.float64(fp_g), **gold)

    # stage-2 NLL and gradients through the reference autograd (LLFlowVQGAN2_arch.py:75-122)
    n2, _, _ = ref_shims.build_reference_stage2(seed=0)
    sd2 = synth.synth_state_dict("netG_stage2", 0)
    n2.load_state_dict(sd2, strict=True)
    n2.train()
    # a parameter with a non-zero bias marks ActNorm as initialised (FlowActNorms.py:36) -> no data-dependent init
    for m in n2.modules():
        if hasattr(m, "inited"):
            m.inited = True
    gt_lat = torch.randn((2, 3, 8, 8), generator=g)
    lr_img = synth.preprocess(torch.rand((2, 3, 32, 32), generator=g))
    import random
    random.seed(1)   # random.random() > train_gt_ratio chooses mean = color_map (LLFlowVQGAN2_arch.py:109)
    while True:
        st = random.getstate()
        if random.random() > 0.2:
            random.setstate(st)
            break
    z2, nll_ref, _ = n2(gt=gt_lat, lr=lr_img, reverse=False)
    nll_ref.mean().backward()
    sd2g = {k: v.clone().requires_grad_(True) for k, v in sd2.items()}
    z2o, nll_o = O.stage2_nll(sd2g, gt_lat, lr_img)
    nll_o.mean().backward()
    log("stage2 nll: ref %s oracle %s ; z maxdiff %.3g" % (nll_ref.tolist(), nll_o.tolist(), maxdiff(z2o, z2)))
    gsel = {}
    named = dict(n2.named_parameters())
    for k in ("flowUpsamplerNet.layers.0.actnorm.bias", "flowUpsamplerNet.layers.0.invconv.weight",
              "flowUpsamplerNet.layers.2.affine.fAffine.0.weight", "flowUpsamplerNet.layers.2.affine.fFeatures.4.logs",
              "flowUpsamplerNet.layers.27.affine.fAffine.4.weight", "RRDB.color_conv.weight", "RRDB.cond_conv.0.weight"):
        gr, go = named[k].grad, sd2g[k].grad
        log("  grad %-55s |g| %.3g  oracle-vs-ref maxdiff %.3g" % (k, float(gr.abs().max()), maxdiff(gr, go)))
        gsel["grad." + k] = gr
    save("stage2", gt_latent=gt_lat, lr=lr_img, z=z2, nll=nll_ref, fingerprint_stage2=np.float64(synth.state_fingerprint(sd2)),
         **gsel)

    # ------------------------------------------------------------------ 3. DCN (deform_conv_cuda_kernel.cu:571-633)
    import torchvision.ops
    C, Co, H, W, dg = 8, 8, 9, 11, 4
    x = torch.randn((2, C, H, W), generator=g)
    off = torch.randn((2, dg * 18, H, W), generator=g) * 2.5
    off[0, :, 0, 0] = 40.0       # far outside
    off[0, :, 0, 1] = -40.0
    off[0, 0::2, 1, 0] = -0.5    # straddles the top border
    off[1, :, 2, 2] = 0.0        # integer positions
    msk = torch.sigmoid(torch.randn((2, dg * 9, H, W), generator=g))
    wt = torch.randn((Co, C, 3, 3), generator=g) * 0.2
    bs = torch.randn((Co,), generator=g)
    y_ref = torchvision.ops.deform_conv2d(x, off, wt, bs, stride=1, padding=1, dilation=1, mask=msk)
    y_o = O.modulated_deform_conv(x, off, msk, wt, bs)
    log("DCN small (C=8, dg=4): oracle(C restatement of the CUDA kernel) vs torchvision maxdiff %.3g" % maxdiff(y_o, y_ref))
    assert maxdiff(y_o, y_ref) < 1e-4
    y0 = O.modulated_deform_conv(x, torch.zeros_like(off), msk, wt, bs)
    # zero offsets: DCN == sum_tap mask_tap * conv_tap  (SURVEY.md section 4 known-answer test)
    ka = bs.view(1, -1, 1, 1).expand(2, Co, H, W).clone()
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    for i in range(3):
        for j in range(3):
            for gi in range(dg):
                cs = slice(gi * (C // dg), (gi + 1) * (C // dg))
                patch = xp[:, cs, i:i + H, j:j + W] * msk[:, gi * 9 + i * 3 + j].unsqueeze(1)
                ka += torch.einsum("bchw,oc->bohw", patch, wt[:, cs, i, j])
    log("DCN zero-offset known answer: maxdiff %.3g" % maxdiff(y0, ka))
    assert maxdiff(y0, ka) < 1e-4
    save("dcn", x=x, offset=off, mask=msk, weight=wt, bias=bs, y=y_ref)

    # ------------------------------------------------------------------ 4. taming blocks (encoder_decoder.py)
    enc = netG.RRDB.encoder
    xb = torch.randn((1, 128, 10, 14), generator=g)
    with torch.no_grad():
        rb_ref = enc.down[0].block[0](xb, None)
        rb_o = O.resnet_block(sd_g, "RRDB.encoder.down.0.block.0", xb)
        xb2 = torch.randn((1, 128, 10, 14), generator=g)
        rb2_ref = enc.down[1].block[0](xb2, None)        # 128 -> 256 with nin_shortcut
        rb2_o = O.resnet_block(sd_g, "RRDB.encoder.down.1.block.0", xb2)
        xa = torch.randn((1, 512, 6, 9), generator=g)
        at_ref = enc.mid.attn_1(xa)
        at_o = O.attn_block(sd_g, "RRDB.encoder.mid.attn_1", xa)
        dn_ref = enc.down[0].downsample(xb)
        dn_o = O.downsample(sd_g, "RRDB.encoder.down.0.downsample", xb)
        up_ref = net_hq.decoder.up[2].upsample(xa)
        up_o = O.upsample(sd_v, "decoder.up.2.upsample", xa)
    log("ResnetBlock 128: %.3g ; ResnetBlock 128->256: %.3g ; AttnBlock 512: %.3g ; Downsample: %.3g ; Upsample: %.3g" %
        (maxdiff(rb_o, rb_ref), maxdiff(rb2_o, rb2_ref), maxdiff(at_o, at_ref), maxdiff(dn_o, dn_ref), maxdiff(up_o, up_ref)))
    save("blocks", x128=xb, res128=rb_ref, x128b=xb2, res128_256=rb2_ref, x512=xa, attn512=at_ref, down128=dn_ref,
         up512=up_ref[:, ::8].contiguous())

    # ------------------------------------------------------------------ 5. full inference pipeline
    for name, (Hh, Ww) in (("pipe_32x48", (32, 48)), ("pipe_64x96", (64, 96))):
        lq, gt = synth.synth_images(1, Hh, Ww, seed=3)
        lr = synth.preprocess(lq)
        cap = {}
        hooks = [netG.RRDB.register_forward_hook(lambda m, i, o: cap.__setitem__("enc", o)),
                 net_hq.quantize.register_forward_hook(lambda m, i, o: cap.__setitem__("vq", (i[0], o))),
                 net_hq.decoder.register_forward_hook(lambda m, i, o: cap.__setitem__("dec", o))]
        with torch.no_grad():
            out_ref, z_ref = netG(net_vq=net_hq, lr=lr, reverse=True)
        for hk in hooks:
            hk.remove()
        st = {}
        out_o = O.glare_infer(sd_g, sd_v, lr, stages=st)
        idx_ref = cap["vq"][1][2][2]
        feats_ref = cap["dec"][1]
        agree = float((st["idx"] == idx_ref).float().mean())
        log("%s: cond_feat %.3g color_map %.3g z_flow %.3g (|z| %.3g) idx agree %.5f vq_feat1 %.3g vq_feat0 %.3g out %.3g (|out| %.3g)" %
            (name, maxdiff(st["cond_feat"], cap["enc"]["cond_feat"]), maxdiff(st["color_map"], cap["enc"]["color_map"]),
             maxdiff(st["z_flow"], z_ref), float(z_ref.abs().max()), agree, maxdiff(st["vq_feat1"], feats_ref[0]),
             maxdiff(st["vq_feat0"], feats_ref[1]), maxdiff(out_o, out_ref), float(out_ref.abs().max())))
        save(name, lq=lq, gt=gt, lr=lr, cond_feat=cap["enc"]["cond_feat"], color_map=cap["enc"]["color_map"],
             mid0=cap["enc"]["mid_feat"][0][:, ::16].contiguous(), mid1=cap["enc"]["mid_feat"][1][:, ::16].contiguous(),
             z_flow=z_ref, idx=idx_ref.to(torch.int32), z_q=cap["vq"][1][0],
             vq_feat1=feats_ref[0][:, ::16].contiguous(), vq_feat0=feats_ref[1][:, ::16].contiguous(), out=out_ref,
             fingerprint_netG=np.float64(fp_g), fingerprint_vqgan=np.float64(fp_v))

    with open(os.path.join(GOLD, "PIN_REPORT.txt"), "w") as f:
        f.write("oracle pinned against the reference's own modules (torch %s)\n" % torch.__version__)
        f.write("\n".join(REPORT) + "\n")


if __name__ == "__main__":
    main()
