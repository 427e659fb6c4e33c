#!/bin/bash
# round 2, call 9: training-step CUDA graph: parity with eager steps, timing through the mirrors (bench alt config)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_zz_flow_train_gpu.py tests/test_flow_gpu.py -m gpu -x -q > gpurun_out/r48_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r48_pytest.log
timeout 900 python - <<'PY' > gpurun_out/r48_train_mirror_timing.txt 2>&1
import sys, time, torch, random
sys.path.insert(0, ".")
from glare_b200 import modules, synth
dev = torch.device("cuda:0")
opt2 = {"train_gt_ratio": 0.2, "datasets": {"train": {"GT_size": 320, "quant": 32}}}
sd_v = synth.synth_state_dict("vqgan", 0)
for use_graph in (True, False):
    netG = modules.VQLLFLOWDeformable(which="netG_stage2", opt=opt2).to(dev)
    netG.load_state_dict(synth.synth_state_dict("netG_stage2", 0), strict=True); netG.train(); netG.train_graph = use_graph
    net_hq = modules.VQModel().to(dev); net_hq.load_state_dict(sd_v, strict=True); net_hq.eval()
    optim = torch.optim.Adam(netG.parameters(), lr=5e-5, betas=(0.9, 0.99))
    gen = torch.Generator().manual_seed(10)
    real_H = torch.rand((4, 3, 320, 320), generator=gen).to(dev)
    var_L = synth.preprocess(torch.rand((4, 3, 320, 320), generator=gen)).to(dev)
    random.seed(10)
    def step():
        optim.zero_grad(set_to_none=True)
        with torch.no_grad():
            enc_gt, _ = net_hq.encode(real_H)
        _, nll, _ = netG(gt=enc_gt.detach(), lr=var_L, reverse=False)
        nll.mean().backward(); optim.step()
        return nll
    for _ in range(6): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): nll = step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("stage-2 step through the mirrors, batch 4 x 320x320, graph=%s: %.1f ms/step (%.2f steps/s, %.1f TFLOP/s algorithmic), nll %s, peak %.1f GB"
          % (use_graph, dt * 1e3, 1 / dt, 12.6 / dt, [round(float(v), 4) for v in nll], torch.cuda.max_memory_allocated() / 2 ** 30))
    del netG, net_hq, optim; torch.cuda.empty_cache()
PY
cat gpurun_out/r48_train_mirror_timing.txt | tail -6
