"""GPU check of the stage-3 training path (run as a child process by tests/test_zz_flow_train_gpu.py; exit code 0 = every check passed).
  1. csrc/loss.cu (MS-SSIM value + gradient) against autograd of the oracle on the CPU, at the training size and at odd small sizes;
  2. the VGG16 perceptual loss on the tensor-core conv path against the oracle;
  3. one whole stage-3 evaluation through the drop-in modules (VQLLFLOWDeformable.train() called like VQLLFLOWD_model.py:207-211, the
     reference's objective on top, total.backward()) against the unmodified reference's numbers (tests/golden/stage3.npz);
  4. timing of that step at train_stage3_LOL.yml's shape (batch 2 x 256 x 256) with Adam."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

FAILS = []


def check(name, ok, detail=""):
    print("%s %s %s" % ("ok  " if ok else "FAIL", name, detail), flush=True)
    if not ok:
        FAILS.append(name)


def images(B, H, W, seed=0, noise=0.15):
    g = torch.Generator().manual_seed(seed)
    gt = torch.nn.functional.avg_pool2d(torch.rand((B, 3, H, W), generator=g), 5, 1, 2)
    return (gt + noise * torch.randn((B, 3, H, W), generator=g)).clamp(0, 1), gt


def main():
    from glare_b200 import losses, modules, synth
    from oracle import losses as OL
    from oracle.gen_golden_stage3 import vgg_state
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")

    # 1. MS-SSIM
    for (B, H, W), normalize, vr in (((2, 256, 256), True, 1), ((1, 52, 67), True, None), ((1, 36, 44), False, None)):
        sr, gt = images(B, H, W, seed=H)
        a, b = sr.clone().to(dev).requires_grad_(True), sr.clone().requires_grad_(True)
        va, vb = losses.msssim(a, gt.to(dev), normalize=normalize, val_range=vr), OL.msssim(b, gt, normalize=normalize, val_range=vr)
        (va * 0.2).backward()
        (vb * 0.2).backward()
        ev, eg = abs(float(va) - float(vb)), float((a.grad.cpu() - b.grad).abs().max()) / float(b.grad.abs().max())
        check("msssim %dx%dx%d normalize=%s" % (B, H, W, normalize), ev <= 2e-6 and eg <= 2e-4, "value diff %.2e grad rel %.2e" % (ev, eg))

    # 2. perceptual.  Gradients through ReLU / max-pool are discontinuous in the activations: a pre-activation within rounding distance of 0
    # takes the other branch on the two devices and moves the input gradient in its receptive field by a whole path's worth (about 10 of
    # 800 K elements per layer here).  So: (a) free-running, relative L2 error with a loose bound; (b) with the CPU pass's activation
    # pattern and pooling choices forced onto the GPU tape, element-wise.
    vsd = vgg_state(0)
    percep = losses.PerceptualNetwork(state_dict=vsd).to(dev)
    sr, gt = images(2, 64, 96, seed=3)
    a, b = sr.clone().to(dev).requires_grad_(True), sr.clone().requires_grad_(True)
    va, vb = percep(a, gt.to(dev)), OL.perceptual(vsd, b, gt)
    va.backward()
    vb.backward()
    d = a.grad.cpu() - b.grad
    ev, e2, em = abs(float(va) - float(vb)) / float(vb), float(d.norm() / b.grad.norm()), float(d.abs().max() / b.grad.abs().max())
    check("perceptual 2x64x96, free-running", ev <= 1e-4 and e2 <= 2e-2 and em <= 0.2, "value rel %.2e grad rel L2 %.2e max %.2e" % (ev, e2, em))
    from encoder_train_emu import TorchLeaves
    from glare_b200 import decoder_train
    leaves = percep.leaves(dev)
    with torch.no_grad():
        gc, gg = losses.VGGGraph(TorchLeaves(), vsd), losses.VGGGraph(leaves, {k: v.to(dev) for k, v in vsd.items()})
        fc, fg = gc.forward(sr), gg.forward(sr.to(dev))
        for sc, sg in zip(gc.switches, gg.switches):
            if sg.dtype == torch.uint8:                   # pooling choice: ATen's flat input index -> the kernel's position code, NHWC
                w_in = 2 * sc.shape[3]
                sc = (((sc // w_in) % 2) * 2 + (sc % w_in) % 2).permute(0, 2, 3, 1).to(torch.uint8)
            sg.copy_(sc.to(dev))
        seeds = [torch.randn(f.shape, generator=torch.Generator().manual_seed(i)) for i, f in enumerate(fc)]
        xc, xg = gc.backward(seeds), gg.backward([t.to(dev) for t in seeds])
    em = float((xc - xg.cpu()).abs().max() / xc.abs().max())
    check("perceptual tape, activation pattern forced", em <= 2e-4, "input gradient rel max %.2e" % em)

    # 3. the reference's stage-3 evaluation
    g = dict(np.load(os.path.join(HERE, "golden", "stage3.npz")))
    netG = modules.VQLLFLOWDeformable().to(dev)
    net_hq = modules.VQModel().to(dev).eval()
    netG.load_state_dict(synth.synth_state_dict("netG", 0), strict=True)
    net_hq.load_state_dict(synth.synth_state_dict("vqgan", 0), strict=True)
    netG.train()
    trainable = [k for k, p in netG.named_parameters() if p.requires_grad]
    check("fix_modules freezes RRDB and the flow", all(k.startswith("deformable_decoder.") for k in trainable) and len(trainable) > 150)
    lq, gt = torch.from_numpy(g["lq"]).to(dev), torch.from_numpy(g["gt"]).to(dev)
    rec, z = netG(net_vq=net_hq, lr=synth.preprocess(lq), reverse=True, reverse_with_grad=True, epses=None, lr_enc=None)
    ez = float((z.cpu() - torch.from_numpy(g["z_flow"])).abs().max())
    check("z_flow", ez < 1e-3, "max abs diff %.2e" % ez)
    erec = float((rec.detach().cpu() - torch.from_numpy(g["rec"])).abs().max())
    check("reconstruction", erec < 1e-3, "max abs diff %.2e" % erec)
    total, terms = losses.stage3_loss(rec, gt, percep)
    for k, v in terms.items():
        d = abs(float(v) - float(g["term." + k]))
        check("term " + k, d <= 1e-4 * max(abs(float(g["term." + k])), 1e-2), "%.8f vs %.8f" % (float(v), float(g["term." + k])))
    total.backward()
    named = dict(netG.named_parameters())
    check("parameters with gradient", {k for k, p in named.items() if p.grad is not None} == set(g["with_grad"].tolist()))
    gmax = max(float(np.abs(g[k]).max()) for k in g if k.startswith("grad."))
    # The offset gradients of the DCN are piecewise constant in the sampling position (bilinear weights): a sampling point within rounding
    # distance of a pixel centre lands in the neighbouring cell on the other device and changes that tap's gradient outright -- measured
    # here: offsets equal to 2e-5 give offset gradients 1e-2 apart in relative L2 (1.4e-4 of the taps flip), which then reaches every
    # parameter upstream at about 5e-3.  Free-running criterion: relative L2 with that bound; the arithmetic itself is checked element-wise
    # in 3b with the sampling cells forced.
    for k in g:
        if k.startswith("grad."):
            ref = torch.from_numpy(g[k])
            d = named[k[5:]].grad.cpu() - ref
            e2 = float(d.norm()) / max(float(ref.norm()), 1e-4 * gmax * ref.numel() ** 0.5)
            em = float(d.abs().max()) / max(float(ref.abs().max()), 1e-4 * gmax)
            check("grad " + k[5:], e2 <= 3e-2 and em <= 0.2, "rel L2 %.2e max %.2e" % (e2, em))
    bad = []
    for k, s in zip(g["with_grad"].tolist(), g["abs_sum"].tolist()):
        got = float(named[k].grad.double().abs().sum())
        if abs(got - s) > 3e-2 * max(s, 1e-4 * gmax * named[k].numel() ** 0.5):
            bad.append((k, got, s))
    check("gradient checksums of all %d parameters" % len(g["with_grad"]), not bad, str(bad[:3]))

    # 3b. the decoder tape, CUDA leaves against torch leaves on the CPU, on the same inputs with the CPU pass's raw offsets / mask logits
    # forced onto the GPU tape before the backward pass: every parameter gradient element-wise
    from oracle import glare_oracle as O
    sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
    with torch.no_grad():
        st = {}
        O.glare_infer(sd_g, sd_v, synth.preprocess(torch.from_numpy(g["lq"])), per_sample_ratio=False, stages=st)
        sd = {k: v for k, v in sd_g.items() if k.startswith("deformable_decoder.")}
        args = (st["z_flow"], [st["vq_feat1"], st["vq_feat0"]], {1: st["mid1"], 0: st["mid0"]})
        tc = decoder_train.DecoderTrainer(TorchLeaves(), sd)
        tg = decoder_train.DecoderTrainer(leaves, {k: v.to(dev) for k, v in sd.items()})
        rc = tc.forward(*args)
        rg = tg.forward(args[0].to(dev), [t.to(dev) for t in args[1]], {k: v.to(dev) for k, v in args[2].items()})
        check("decoder tape forward", float((rc - rg.cpu()).abs().max()) <= 1e-4 * float(rc.abs().max()))
        for ic, ig in zip(tc.offset_ids, tg.offset_ids):
            tg.vals[ig].copy_(tc.vals[ic].to(dev))
        seed = torch.randn(rc.shape, generator=torch.Generator().manual_seed(0))
        gcpu, ggpu = tc.backward(seed), tg.backward(seed.to(dev))
    gmax = max(float(v.abs().max()) for v in gcpu.values())
    worst, wk = 0.0, None
    for k, v in gcpu.items():
        rel = float((v - ggpu[k].cpu()).abs().max()) / max(float(v.abs().max()), 1e-4 * gmax)      # k.bias: exactly 0 in theory
        if rel > worst:
            worst, wk = rel, k
    check("decoder tape backward, sampling cells forced: all %d parameters" % len(gcpu), set(gcpu) == set(ggpu) and worst <= 1e-3,
          "worst rel max %.2e (%s)" % (worst, wk))

    # 4. timing at the training shape
    B, S = 2, 256
    lq, gt = synth.synth_images(B, S, S, seed=1)
    lr, gt = synth.preprocess(lq).to(dev), gt.to(dev)
    opt = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=5e-5, betas=(0.9, 0.99))

    def step():
        opt.zero_grad(set_to_none=True)
        rec, _ = netG(net_vq=net_hq, lr=lr, reverse=True, reverse_with_grad=True)
        total, _ = losses.stage3_loss(rec, gt, percep)
        total.backward()
        opt.step()
        return total

    for _ in range(3):
        first = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    n = 10
    for _ in range(n):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("stage-3 step, batch %d x %dx%d through the drop-in modules + Adam: %.1f ms (%.1f ms wall), objective %.5f -> %.5f, peak memory %.2f GB"
          % (B, S, S, ms, (time.time() - t0) * 1e3 / n, float(first), float(last), torch.cuda.max_memory_allocated() / 2 ** 30), flush=True)
    check("objective finite and decreasing under Adam", bool(torch.isfinite(last)) and float(last) < float(first))
    print("FAILED: %s" % FAILS if FAILS else "ALL OK")
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())
