"""BASELINE config 4 shape (stage-2 flow training: batch 4, 320x320 crops, latent 80x80) through glare_b200.encoder_train.stage2_step:
wall-clock per step and the split encoder forward / flow forward+backward / encoder backward.  python tools/gpu/train_probe.py [steps]
(tensor-core weight gradient by default; GLARE_WGRAD_FMA=1 selects the fp32 split-K GEMM)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from glare_b200 import encoder_train, flow, flow_train, synth  # noqa: E402
from glare_b200.dense import make_dense  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
sd = {k: v.to(dev) for k, v in synth.synth_state_dict("netG_stage2", 0).items()}
plan = flow.FlowPlan({k: v.cpu() for k, v in sd.items()}, dev)
import os  # noqa: E402
dense = make_dense(os.environ.get("GLARE_DENSE", "auto"))          # GLARE_DENSE=tc-bf16: the bf16 training configuration
leaves = encoder_train.CudaLeaves(dense)
fkern = flow_train.CudaKernels(mode=dense.mode if dense.mode in (0, 4) else 4)
conv = lambda x, w: dense.conv2d(x, w).float()      # noqa: E731
gen = torch.Generator().manual_seed(10)             # train_stage2_LOL.yml manual_seed
lr = synth.preprocess(torch.rand((4, 3, 320, 320), generator=gen)).to(dev)
gt = torch.randn((4, 3, 80, 80), generator=gen).to(dev)


def step():
    t = [time.perf_counter()]
    enc = encoder_train.EncoderTrainer(leaves, sd)
    heads = enc.forward(lr)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    nll, _, _, g_ft, g_mean, grads = flow_train.nll_forward_backward(plan, sd, gt, heads["cond_feat"], heads["color_map"], conv, kernels=fkern)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    grads.update(enc.backward(g_ft, g_mean))
    torch.cuda.synchronize(); t.append(time.perf_counter())
    step.grads = grads
    return nll, [b - a for a, b in zip(t, t[1:])]


with torch.no_grad():
    step()
    acc = [0.0, 0.0, 0.0]
    for _ in range(steps):
        nll, dt = step()
        acc = [a + b for a, b in zip(acc, dt)]
print("stage-2 step, batch 4 x 320x320 (latent 80x80), dense %s: nll %s" % (dense.dtype_name, [round(float(x), 4) for x in nll]))
if os.environ.get("GLARE_GRAD_DUMP"):
    torch.save({k: v.detach().cpu() for k, v in step.grads.items()}, os.environ["GLARE_GRAD_DUMP"])
if os.environ.get("GLARE_GRAD_COMPARE"):
    ref = torch.load(os.environ["GLARE_GRAD_COMPARE"])
    rel = sorted(float((step.grads[k].cpu() - v).norm() / max(float(v.norm()), 1e-12)) for k, v in ref.items() if float(v.norm()) > 0)
    cos = sorted(float(torch.nn.functional.cosine_similarity(step.grads[k].cpu().flatten(), v.flatten(), dim=0)) for k, v in ref.items()
                 if float(v.norm()) > 0)
    print("  gradients against %s: relative L2 error per tensor median %.2e, 90th percentile %.2e, max %.2e; cosine median %.5f, min %.5f"
          % (os.environ["GLARE_GRAD_COMPARE"], rel[len(rel) // 2], rel[len(rel) * 9 // 10], rel[-1], cos[len(cos) // 2], cos[0]))
print("  encoder forward %.1f ms | flow forward + backward %.1f ms | encoder backward %.1f ms | total %.1f ms  (peak memory %.1f GB)"
      % tuple([1e3 * a / steps for a in acc] + [1e3 * sum(acc) / steps, torch.cuda.max_memory_allocated() / 2 ** 30]))

if len(sys.argv) > 2 and sys.argv[2] == "--profile":
    # kernel-time breakdown of one step (torch.profiler / CUPTI): which kernels the 0.2 s go to
    from torch.profiler import ProfilerActivity, profile
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
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
    for name, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:28]:
        print("  %-62s %5d launches %9.2f ms %5.1f %%" % (name, n, t / 1e3, 100 * t / tot))
