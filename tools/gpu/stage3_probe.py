"""train_stage3_LOL.yml shape (batch 2 x 256x256) through the drop-in mirrors: time per phase of one stage-3 step and the kernel-time
breakdown (torch.profiler / CUPTI).   python tools/gpu/stage3_probe.py [steps] [--profile]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from glare_b200 import losses, modules, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 5
dev = torch.device("cuda:0")
netG = modules.VQLLFLOWDeformable().to(dev)
net_hq = modules.VQModel().to(dev).eval()
netG.load_state_dict(synth.synth_state_dict("netG", 0), strict=True)
net_hq.load_state_dict(synth.synth_state_dict("vqgan", 0), strict=True)
netG.train()
g = torch.Generator().manual_seed(4000)
percep = losses.PerceptualNetwork(state_dict={"%d.%s" % (i, n): (torch.randn((co, ci, 3, 3), generator=g) * (2.0 / (9 * ci)) ** 0.5 if n == "weight"
                                                                else torch.zeros(co)) for i, ci, co in losses.VGG_CONVS for n in ("weight", "bias")}).to(dev)
opt = torch.optim.Adam([p for p in netG.parameters() if p.requires_grad], lr=5e-5, betas=(0.9, 0.99))
lq, gt = synth.synth_images(2, 256, 256, seed=1)
lr, gt = synth.preprocess(lq).to(dev), gt.to(dev)


def step(sync=False):
    t = [time.perf_counter()]

    def mark():
        if sync:
            torch.cuda.synchronize()
            t.append(time.perf_counter())

    opt.zero_grad(set_to_none=True)
    rec, _ = netG(net_vq=net_hq, lr=lr, reverse=True, reverse_with_grad=True)
    mark()
    total, _ = losses.stage3_loss(rec, gt, percep)
    mark()
    total.backward()
    mark()
    opt.step()
    mark()
    return total, [b - a for a, b in zip(t, t[1:])]


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    total, _ = step()
e1.record()
torch.cuda.synchronize()
print("stage-3 step, batch 2 x 256x256: %.1f ms per step on the device (%.1f ms wall), objective %.5f, peak memory %.1f GB"
      % (e0.elapsed_time(e1) / steps, 1e3 * (time.perf_counter() - t0) / steps, float(total), torch.cuda.max_memory_allocated() / 2 ** 30))
acc = [0.0] * 4
for _ in range(steps):
    _, dt = step(sync=True)
    acc = [a + b for a, b in zip(acc, dt)]
print("  frozen stages + decoder forward %.1f ms | losses forward %.1f ms | backward (losses + decoder) %.1f ms | Adam %.1f ms"
      % tuple(1e3 * a / steps for a in acc))
if "--profile" in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
            name = ev.name.split("<")[0].split("(")[0][:60]
            r = rows.setdefault(name, [0.0, 0])
            r[0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            r[1] += 1
    tot = sum(v[0] for v in rows.values())
    print("kernel-time breakdown of one step: %.1f ms of device time in %d kernels" % (tot / 1e3, sum(v[1] for v in rows.values())))
    for name, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:30]:
        print("  %-62s %5d launches %9.2f ms %5.1f %%" % (name, n, t / 1e3, 100 * t / tot))
