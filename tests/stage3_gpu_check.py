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

    # 2. perceptual
    vsd = vgg_state(0)
    percep = losses.PerceptualNetwork(state_dict=vsd).to(dev)
    sr, gt = images(2, 64, 96, seed=3)
    a, b = sr.clone().to(dev).requires_grad_(True), sr.clone().requires_grad_(True)
    va, vb = percep(a, gt.to(dev)), OL.perceptual(vsd, b, gt)
    va.backward()
    vb.backward()
    ev, eg = abs(float(va) - float(vb)) / float(vb), float((a.grad.cpu() - b.grad).abs().max()) / float(b.grad.abs().max())
    check("perceptual 2x64x96", ev <= 1e-4 and eg <= 1e-3, "value rel %.2e grad rel %.2e" % (ev, eg))

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
    check("z_flow", float((z.cpu() - torch.from_numpy(g["z_flow"])).abs().max()) < 1e-4)
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
    worst = 0.0
    for k in g:
        if k.startswith("grad."):
            ref = torch.from_numpy(g[k])
            rel = float((named[k[5:]].grad.cpu() - ref).abs().max()) / max(float(ref.abs().max()), 1e-4 * gmax)
            worst = max(worst, rel)
            check("grad " + k[5:], rel <= 5e-3, "rel %.2e" % rel)
    bad = []
    for k, s in zip(g["with_grad"].tolist(), g["abs_sum"].tolist()):
        got = float(named[k].grad.double().abs().sum())
        if abs(got - s) > 1e-2 * max(s, 1e-4 * gmax * named[k].numel() ** 0.5):
            bad.append((k, got, s))
    check("gradient checksums of all %d parameters" % len(g["with_grad"]), not bad, str(bad[:3]))

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
