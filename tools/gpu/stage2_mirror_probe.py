"""per-step device time of the stage-2 training step through the drop-in mirrors (the sequence bench.py's alt config times)"""
import random
import sys
import time

import torch

sys.path.insert(0, ".")
from glare_b200 import modules, synth  # noqa: E402

dev = torch.device("cuda:0")
opt2 = {"train_gt_ratio": 0.2, "datasets": {"train": {"GT_size": 320, "quant": 32}}}
netG = modules.VQLLFLOWDeformable(which="netG_stage2", opt=opt2).to(dev)
netG.load_state_dict(synth.synth_state_dict("netG_stage2", 0), strict=True)
netG.train()
net_hq = modules.VQModel().to(dev)
net_hq.load_state_dict(synth.synth_state_dict("vqgan", 0), strict=True)
net_hq.eval()
named = [(k, p) for k, p in netG.named_parameters()]
optim = torch.optim.Adam([p for _, p in named], lr=5e-5, betas=(0.9, 0.99))
gen = torch.Generator().manual_seed(10)
real_H = torch.rand((4, 3, 320, 320), generator=gen).to(dev)
var_L = synth.preprocess(torch.rand((4, 3, 320, 320), generator=gen)).to(dev)
random.seed(10)
for i in range(14):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    optim.zero_grad(set_to_none=True)
    with torch.no_grad():
        encoder_gt, _ = net_hq.encode(real_H)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    _, nll, _ = netG(gt=encoder_gt.detach(), lr=var_L, reverse=False)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    nll.mean().backward()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    optim.step()
    torch.cuda.synchronize(); t.append(time.perf_counter())
    print("step %2d: vqgan encode %.1f | objective + gradients %.1f | backward() %.1f | Adam %.1f | total %.1f ms   mem %.1f GB" %
          tuple([i] + [1e3 * (b - a) for a, b in zip(t, t[1:])] + [1e3 * (t[-1] - t[0]), torch.cuda.memory_allocated() / 2 ** 30]), flush=True)

# the same step without host synchronisation (what bench.py times), one event per step
import gc  # noqa: E402


def step():
    optim.zero_grad(set_to_none=True)
    with torch.no_grad():
        encoder_gt, _ = net_hq.encode(real_H)
    _, nll, _ = netG(gt=encoder_gt.detach(), lr=var_L, reverse=False)
    nll.mean().backward()
    optim.step()


for label, before in (("as is", lambda: None), ("gc disabled", gc.disable)):
    before()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(12):
        step()
        evs[i + 1].record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("unsynchronised, %s: per-step device ms %s | host issue time %.1f ms per step" %
          (label, [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(12)], 1e3 * host / 12), flush=True)
