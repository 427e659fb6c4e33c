import random, sys, time, torch
sys.path.insert(0, ".")
from glare_b200 import modules, synth
dev = torch.device("cuda:0")
opt2 = {"train_gt_ratio": 0.0, "datasets": {"train": {"GT_size": 320, "quant": 32}}}
for use_graph in (False, True):
    netG = modules.VQLLFLOWDeformable(which="netG_stage2", opt=opt2).to(dev)
    netG.load_state_dict(synth.synth_state_dict("netG_stage2", 0), strict=True)
    netG.train(); netG.train_graph = use_graph
    optim = torch.optim.Adam(netG.parameters(), lr=5e-5, betas=(0.9, 0.99))
    gen = torch.Generator().manual_seed(10)
    gt = torch.randn((4, 3, 80, 80), generator=gen).to(dev)
    var_L = synth.preprocess(torch.rand((4, 3, 320, 320), generator=gen)).to(dev)
    def step():
        optim.zero_grad(set_to_none=True)
        _, nll, _ = netG(gt=gt, lr=var_L, reverse=False)
        nll.mean().backward()
        optim.step()
    for _ in range(4): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print("train_graph=%s: %.1f ms per step (objective + gradients + Adam, no VQGAN encode), mem %.1f GB" % (use_graph, e0.elapsed_time(e1) / 10, torch.cuda.max_memory_allocated() / 2**30))
    del netG, optim
