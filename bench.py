#!/usr/bin/env python
"""Throughput benchmark of the GLARE hot path (BASELINE.json metric: 600x400 images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl glare|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One step = one pass of the whole inference path (cond-encoder -> inverse flow -> VQ lookup -> VQGAN decoder
-> AFT/DCN decoder) over one batch of 15 synthetic LOL-eval-shape images (400x600, reflect-padded to 420x620
as infer_dataset_lol.py:124 does) per GPU -- BASELINE.json configs[1].  Images shard across ranks with no
data-path collective; the only NCCL call is the final all_gather of the outputs (weak scaling).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = through GlareEnhancer.enhance with
pinned host uint8 buffers (H2D + pre/post-processing + D2H inside the timed region), `roofline` = the dominant
kernel of libglare_b200.so timed live with CUDA events, `cpu_baseline` = the CPU oracle port on the host cores.
`--impl reference` times the reference algorithm's CPU path (the oracle port: the reference is Python and
/root/reference does not exist on the GPU box) with all host threads.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "600x400 images/sec"          # BASELINE.json metric; the workload (LOL eval15 shape, batch 15 per GPU) is in config
H, W, BATCH = 400, 600, 15


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tensor": p["bf16_tflops_sustained"], "tensor_burst": p["bf16_tflops"], "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor": 1400.0, "tensor_burst": 1590.0, "src": "fallback"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(batch, seed):
    from glare_b200 import synth
    lq, gt = synth.synth_images(batch, H, W, seed=seed)
    return lq, gt


def cpu_oracle_step(sd_g, sd_v, lr_one):
    from oracle import glare_oracle as O
    t0 = time.perf_counter()
    O.glare_infer(sd_g, sd_v, lr_one)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """CPU arm: the reference algorithm (oracle port) on the host cores, rank 0 only."""
    if rank != 0:
        return
    from glare_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
    lq, _ = synth_batch(1, 0)
    lr = synth.preprocess(synth.pad_lol(lq))
    budget = float(os.environ.get("GLARE_BENCH_CPU_BUDGET_S", "200"))
    t_start = time.perf_counter()
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_step(sd_g, sd_v, lr)
    times = []
    while len(times) < args.steps and (not times or time.perf_counter() - t_start + times[-1] < budget):
        times.append(cpu_oracle_step(sd_g, sd_v, lr))
    ips = len(times) / sum(times)
    sample = "1 image (420x620 padded) of the 15-image batch per step; %d of %d requested steps inside a %.0f s budget" % (
        len(times), args.steps, budget)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "LOL eval15-shape inference 600x400 (padded 420x620), fp32, CPU oracle port of the reference path",
                       "batch_per_step": 1},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def alt_configs(rank, world, dev, sd_g, sd_v, barrier, flush, fp32_engine, quick=False):
    """The other single-box configurations of BASELINE.json, measured with the same rules (device events, barrier on both sides, max over
    ranks, L2 flush between iterations, weak scaling): configs[2] bf16 operands at 8 images per GPU (bs 64 over 8 GPUs) and configs[4]
    1920x1080 unpaired inference (auto_padding to 1088x1936) at 1 / 2 / 4 images per GPU.  Returned as the `alt_configs` object of the line."""
    import torch.distributed as dist
    from glare_b200 import synth
    from glare_b200.api import GlareEnhancer
    from glare_b200.dense import make_dense
    psnr = lambda a, b: float(10 * torch.log10(1.0 / torch.mean((a - b) ** 2)))       # noqa: E731  (utils/utils2.py:32-36 on [0,1] images)

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
            flush.zero_()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    def timed_each(fn, steps, warm):
        """like `timed`, with one event per step: -> (mean ms per step over the ranks' max, {median, max} of this rank's steps) -- the
        training steps run thousands of small launches from Python, and a collector pause or an allocator refill in one step moves a 8-step mean"""
        for _ in range(warm):
            fn()
            flush.zero_()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        gc.collect()
        gc.freeze()                       # the interpreter's old generation (millions of objects of torch and this process) is not re-walked by
        barrier()                         # a full collection that happens to start inside a timed step; the steps' own garbage still is
        evs[0].record()
        for i in range(steps):
            fn()
            flush.zero_()
            evs[i + 1].record()
        barrier()
        gc.unfreeze()
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
        t = torch.tensor([evs[0].elapsed_time(evs[steps])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, {"ms_per_step_median": per[len(per) // 2], "ms_per_step_max": per[-1]}

    out = {}
    pk = peaks()
    # ---- configs[2]: bf16 tensor-core operands (fp32 accumulate, fp32 GroupNorm / softmax / residual stream), 8 images per GPU
    B = 8
    lq, gt = synth_batch(B, seed=100 + rank)
    enh = GlareEnhancer(sd_g, sd_v, device=dev, pad="lol", dense=make_dense("tc-bf16"))
    host_u8 = (lq.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
    host_out = torch.empty_like(host_u8).pin_memory()
    dev8 = host_u8.to(dev)
    lr_dev, box = enh.preprocess(dev8)
    steps, warm = (3, 3) if quick else (10, 3)
    ms = timed(lambda: enh.enhance_device(dev8), steps, warm)         # one CUDA-graph replay per step, batch resident in HBM
    ms_e2e = timed(lambda: enh.enhance(host_u8, out=host_out), steps, 2)
    crop = lambda o: o[:, :, box[0]:box[1], box[2]:box[3]].clamp(0, 1).float().cpu()       # noqa: E731
    o16, o32 = crop(enh.engine.infer(lr_dev)), crop(fp32_engine.infer(lr_dev))
    gtq = gt                                                                               # clean target of the synthetic pair
    ips = world * B / (ms / 1e3)
    out["lolv2_real_bf16_bs64_over_8gpus"] = {
        "workload": "configs[2]: 8 images 600x400 (padded 420x620) per GPU per step, bf16 tensor-core operands", "dtype": "bf16",
        "value": ips, "unit": "images/s", "ms_per_step": ms, "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "images/s",
                                                                  "h2d_bytes_per_step": host_u8.numel(), "d2h_bytes_per_step": host_out.numel()},
        "frac_of_bf16_ceiling": ips / world / (pk["tensor"] / 13.16),
        "ceiling_images_per_s_per_gpu": pk["tensor"] / 13.16,
        "dpsnr_vs_fp32_path_db": abs(psnr(o16, gtq) - psnr(o32, gtq)), "pixel_mean_abs_diff_vs_fp32_path": float((o16 - o32).abs().mean())}
    del enh, lr_dev
    gc.collect()                      # engines hold reference cycles (graphs <-> closures): without this their CUDA graphs and pools are torn
    torch.cuda.empty_cache()          # down by the cyclic collector at a random moment inside the NEXT timed region (seen: +0.5 s in one step)
    # ---- configs[3]: one stage-2 training step through the drop-in mirrors, the call sequence of LLFlow_model.optimize_parameters
    # (LLFlow_model.py:181-232): frozen VQGAN encodes the ground truth, netG(gt=..., lr=..., reverse=False) -> nll.mean().backward()
    # (objective + every gradient from libglare_b200.so), gradient all-reduce over the ranks, Adam step.  Batch 4 x 320x320 per GPU.
    if not quick:
        from glare_b200 import modules
        from glare_b200.parallel import allreduce_gradients
        opt2 = {"train_gt_ratio": 0.2, "datasets": {"train": {"GT_size": 320, "quant": 32}}}          # train_stage2_LOL.yml
        netG = modules.VQLLFLOWDeformable(which="netG_stage2", opt=opt2).to(dev)
        netG.load_state_dict(synth.synth_state_dict("netG_stage2", 0), strict=True)
        netG.train()
        net_hq = modules.VQModel().to(dev)
        net_hq.load_state_dict(sd_v, strict=True)
        net_hq.eval()
        named = [(k, p) for k, p in netG.named_parameters()]
        optim = torch.optim.Adam([p for _, p in named], lr=5e-5, betas=(0.9, 0.99))
        gen = torch.Generator().manual_seed(10 + rank)                                                # train_stage2_LOL.yml manual_seed
        real_H = torch.rand((4, 3, 320, 320), generator=gen).to(dev)
        var_L = synth.preprocess(torch.rand((4, 3, 320, 320), generator=gen)).to(dev)
        import random
        random.seed(10)
        last = {}

        def train_step():
            optim.zero_grad(set_to_none=True)
            with torch.no_grad():
                encoder_gt, _ = net_hq.encode(real_H)
            _, nll, _ = netG(gt=encoder_gt.detach(), lr=var_L, reverse=False)
            nll.mean().backward()
            if world > 1:
                grads = allreduce_gradients({k: p.grad for k, p in named if p.grad is not None})
                for k, p in named:
                    if p.grad is not None:
                        p.grad = grads[k].to(p.grad.dtype).reshape(p.grad.shape)
            optim.step()
            last["nll"] = nll.detach()

        ms_t, spread_t = timed_each(train_step, 8, 4)    # (the caching allocator needs a few steps to settle after the inference configs)
        out["stage2_training_step"] = {
            "workload": "configs[3]: one stage-2 step (frozen-VQGAN encode of GT, flow NLL forward + backward, Adam) through the drop-in "
                        "mirrors, batch 4 x 320x320 per GPU, fp32-grade tensor-core operands (bf16x3), train_gt_ratio 0.2",
            "value": world * 4 / (ms_t / 1e3), "unit": "samples/s", "ms_per_step": ms_t, **spread_t, "steps_per_s_per_gpu": 1e3 / ms_t,
            "algorithmic_tflop_per_step": 12.6, "achieved_tflops_per_gpu": 12.6 / (ms_t / 1e3), "frac_of_bf16_peak": 12.6 / (ms_t / 1e3) / pk["tensor"],
            "nll_last": [round(float(v), 4) for v in last["nll"].cpu()],
            "collective": "one flat fp32 gradient all-reduce (NCCL)" if world > 1 else None}
        # the same step with bf16 tensor-core operands (the precision configs[3] names): convolutions, attention GEMMs and weight gradients on
        # single-piece bf16 operands, fp32 accumulation, fp32 memory-bound kernels and optimizer
        netG.dense_name, netG._train_ctx = "tc-bf16", None
        ms_tb, spread_tb = timed_each(train_step, 8, 3)
        out["stage2_training_step_bf16"] = {
            "workload": "configs[3] at its named precision: the same stage-2 step with bf16 tensor-core operands", "dtype": "bf16",
            "value": world * 4 / (ms_tb / 1e3), "unit": "samples/s", "ms_per_step": ms_tb, **spread_tb,
            "nll_last": [round(float(v), 4) for v in last["nll"].cpu()]}
        del netG, optim
        gc.collect()
        torch.cuda.empty_cache()
        # ---- stage 3 (train_stage3_LOL.yml: batch 2 x 256x256): the call sequence of VQLLFLOWDModel.optimize_parameters
        # (VQLLFLOWD_model.py:187-232): netG(net_vq=..., lr=..., reverse=True, reverse_with_grad=True) with encoder / flow / VQGAN frozen,
        # |sr - gt| + 0.01 VGG16 perceptual + 0.2 (1 - MS-SSIM), total.backward() (decoder tape with the DCN backward kernels, csrc/loss.cu),
        # gradient all-reduce, Adam on deformable_decoder.*
        from glare_b200 import losses
        netG3 = modules.VQLLFLOWDeformable().to(dev)
        netG3.load_state_dict(sd_g, strict=True)
        netG3.train()
        vgen = torch.Generator().manual_seed(4000)
        percep = losses.PerceptualNetwork(state_dict={"%d.%s" % (i, n): (torch.randn((co, ci, 3, 3), generator=vgen) * (2.0 / (9 * ci)) ** 0.5
                                                                        if n == "weight" else torch.zeros(co))
                                                      for i, ci, co in losses.VGG_CONVS for n in ("weight", "bias")}).to(dev)
        named3 = [(k, p) for k, p in netG3.named_parameters() if p.requires_grad]
        optim3 = torch.optim.Adam([p for _, p in named3], lr=5e-5, betas=(0.9, 0.99))
        lq3, gt3 = synth.synth_images(2, 256, 256, seed=400 + rank)
        var_L3, real_H3 = synth.preprocess(lq3).to(dev), gt3.to(dev)
        last3 = {}

        def train_step3():
            optim3.zero_grad(set_to_none=True)
            rec, _ = netG3(net_vq=net_hq, lr=var_L3, reverse=True, reverse_with_grad=True)
            total, terms = losses.stage3_loss(rec, real_H3, percep)
            total.backward()
            if world > 1:
                grads = allreduce_gradients({k: p.grad for k, p in named3 if p.grad is not None})
                for k, p in named3:
                    if p.grad is not None:
                        p.grad = grads[k].to(p.grad.dtype).reshape(p.grad.shape)
            optim3.step()
            last3["total"] = total.detach()

        ms_3, spread_3 = timed_each(train_step3, 8, 4)
        out["stage3_training_step"] = {
            "workload": "one stage-3 step (train_stage3_LOL.yml shape: batch 2 x 256x256 per GPU) through the drop-in mirrors: frozen encoder / "
                        "flow / VQGAN forward, deformable-decoder forward + backward (DCN backward in the loop), L1 + VGG16 perceptual (synthetic "
                        "weights) + MS-SSIM objective, Adam; fp32-grade tensor-core operands (bf16x3)",
            "value": world * 2 / (ms_3 / 1e3), "unit": "samples/s", "ms_per_step": ms_3, **spread_3, "objective_last": round(float(last3["total"]), 5),
            "collective": "one flat fp32 gradient all-reduce (NCCL)" if world > 1 else None}
        del netG3, net_hq, optim3, percep
        gc.collect()
        torch.cuda.empty_cache()
    # ---- single-image latency of the bench shape (infer_dataset_lol.py / infer_unpaired.py run batch 1): host uint8 in -> host uint8 out
    enh1 = GlareEnhancer(sd_g, sd_v, device=dev, pad="lol", dense=fp32_engine.dense)
    lq1, _ = synth_batch(1, seed=300 + rank)
    h1 = (lq1.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
    o1 = torch.empty_like(h1).pin_memory()
    ms1 = timed(lambda: enh1.enhance(h1, out=o1), 5 if quick else 20, 3)
    out["latency_600x400_batch1"] = {"workload": "one 600x400 image end to end through GlareEnhancer.enhance (H2D, one CUDA-graph replay, D2H), fp32-grade",
                                     "value": ms1, "unit": "ms/image", "images_per_s": world * 1e3 / ms1}
    del enh1
    # ---- configs[4]: 1920x1080, the fp32-grade default backend, batch swept
    enh = GlareEnhancer(sd_g, sd_v, device=dev, pad="auto", dense=fp32_engine.dense)
    sweep = {}
    for Bh in ((1,) if quick else (1, 2, 4)):
        lqh, _ = synth.synth_images(Bh, 1080, 1920, seed=200 + rank)
        hu8 = (lqh.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
        hout = torch.empty_like(hu8).pin_memory()
        ms = timed(lambda: enh.enhance(hu8, out=hout), 2, 1)             # through the public API: host buffers in and out
        sweep["batch_%d" % Bh] = {"value": world * Bh / (ms / 1e3), "unit": "images/s", "ms_per_step": ms,
                                  "frac_of_bf16_ceiling": Bh / (ms / 1e3) / (pk["tensor"] / 447.0)}
    out["unpaired_1080p_fp32"] = {"workload": "configs[4]: 1920x1080 images (auto_padding to 1088x1936, 131 648 latent tokens), end to end "
                                              "through GlareEnhancer.enhance with host buffers, fp32-grade operands (bf16x3)",
                                  "dtype": fp32_engine.dense.dtype_name, "ceiling_images_per_s_per_gpu_bf16": pk["tensor"] / 447.0,
                                  "per_gpu_batch_sweep": sweep}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="glare", choices=["glare", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--dense", default=os.environ.get("GLARE_DENSE", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--no-alt", action="store_true", help="skip the alt_configs measurements (bf16 config 3, 1080p config 5)")
    ap.add_argument("--quick-alt", action="store_true", help="alt_configs with fewer steps and 1080p at batch 1 only")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from glare_b200 import ops, synth
    from glare_b200.api import GlareEnhancer
    from glare_b200.dense import make_dense

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
    dense = make_dense(args.dense)
    enh = GlareEnhancer(sd_g, sd_v, device=dev, pad="lol", dense=dense)
    eng = enh.engine
    B = args.batch
    use_graph = not args.no_graph
    lq, gt = synth_batch(B, seed=rank)
    host_u8 = (lq.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
    host_out = torch.empty_like(host_u8).pin_memory()
    dev_u8 = host_u8.to(dev)                                               # `value`: the batch is resident in HBM (uint8 NHWC, as decoded)
    lr_dev, box = enh.preprocess(dev_u8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    from glare_b200.parallel import AsyncGather
    gather = AsyncGather()                                                 # world > 1: all_gather of the uint8 results on a side stream

    def step(graph=use_graph):
        # pre-processing kernel -> the whole network -> post-processing kernel; one CUDA-graph replay per step when graph is on
        gather.submit(enh.enhance_device(dev_u8, graph=graph))

    for _ in range(max(args.warmup, 3)):
        step()
        flush.zero_()
    warm = max(args.warmup, 3)

    # ---- device-resident throughput (`value`)
    n0 = ops.LAUNCHES
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
        flush.zero_()
    gather.wait()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ops.LAUNCHES - n0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- per-kernel timers for the roofline: the same step, launched eagerly with a CUDA-event bracket around every operator (a captured
    # graph has no per-kernel brackets); rank 0 only needs it, every rank runs it to stay in step
    eng.timers = {}
    if hasattr(dense, "timers"):
        dense.timers = {}
    inst_steps = min(args.steps, 5)
    e0.record()
    for _ in range(inst_steps):
        step(graph=False)
        flush.zero_()
    gather.wait()
    e1.record()
    barrier()
    inst_ms = e0.elapsed_time(e1) / inst_steps
    timers, eng.timers = eng.timers, None
    if hasattr(dense, "timers"):
        dense.last_timers, dense.last_steps, dense.timers = dense.timers, inst_steps, None

    # ---- end to end through the public API with host buffers
    for _ in range(2):
        enh.enhance(host_u8, out=host_out, graph=use_graph)
    barrier()
    e0.record()
    for _ in range(args.steps):
        enh.enhance(host_u8, out=host_out, graph=use_graph)
        flush.zero_()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e = {"value": world * B * args.steps / (e2e_ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": host_u8.numel(),
           "d2h_bytes_per_step": host_out.numel()}

    alt = None
    if not args.no_alt and args.dense in ("auto", "tc-bf16x3"):
        try:
            alt = alt_configs(rank, world, dev, sd_g, sd_v, barrier, flush, eng, quick=args.quick_alt)
        except Exception as exc:                       # the headline line must survive a failure in the side measurements
            import traceback
            alt = {"error": "%s: %s" % (type(exc).__name__, exc), "traceback": traceback.format_exc()[-1500:]}

    if rank == 0:
        pk = peaks()
        roof = dense.roofline(timers, eng, B, lr_dev.shape, pk)
        if roof is not None and "share_of_step" in roof:
            roof["share_of_step"] = roof["ms_per_step"] / inst_ms
            roof["timing"] = ("CUDA-event brackets around every launch of the kernel in %d eagerly launched steps (%.1f ms/step) run right "
                              "after the graph-replayed timed region" % (inst_steps, inst_ms))
        cpu = None
        if not args.no_cpu_baseline and world == 1:        # reported on rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            one = lr_dev[:1].cpu()
            dt = cpu_oracle_step(sd_g, sd_v, one)
            cpu = {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "1 image (420x620 padded) of the 15-image batch, CPU oracle port, %d torch threads, %.1f s" % (cores, dt)}
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": dense.dtype_name, "data": "synthetic",
                "config": {"workload": "LOL eval15-shape batch inference: 15 images 600x400 (reflect-padded to 420x620) per GPU per step",
                           "batch_per_gpu": B, "dense_backend": dense.name,
                           "library_fallbacks_per_run": getattr(dense, "fallbacks", None), "parallelism": "images sharded, dp%d" % world,
                           "launch": ("one CUDA graph replay per step (pre-processing + network + post-processing), %d kernels per graph"
                                      % (launches // max(1, args.steps))) if use_graph else "eager (one ctypes call per kernel)",
                           "gather": "uint8 results all_gathered on a side stream (NCCL), overlapped with the next step" if world > 1 else None,
                           "l2": "256 MiB buffer written between timed iterations (L2 flush); activations per step exceed L2"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "peaks": pk["src"], "alt_configs": alt,
                "breakdown_ms_per_step": dense.breakdown(timers, inst_steps) if hasattr(dense, "breakdown") else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
