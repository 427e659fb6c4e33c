"""edge-by-edge comparison of the stage-3 decoder tape, CUDA leaves against torch leaves on the CPU, on the inputs of tests/golden/stage3.npz --
debugging aid for tests/stage3_gpu_check.py.   python tools/gpu/stage3_debug.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from encoder_train_emu import TorchLeaves  # noqa: E402
from glare_b200 import decoder_train, encoder_train, synth  # noqa: E402
from glare_b200.dense import make_dense  # noqa: E402
from oracle import glare_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
g = dict(np.load("tests/golden/stage3.npz"))
sd_g, sd_v = synth.synth_state_dict("netG", 0), synth.synth_state_dict("vqgan", 0)
with torch.no_grad():
    st = {}
    O.glare_infer(sd_g, sd_v, synth.preprocess(torch.from_numpy(g["lq"])), per_sample_ratio=False, stages=st)
sd = {k: v for k, v in sd_g.items() if k.startswith("deformable_decoder.")}
args = (st["z_flow"], [st["vq_feat1"], st["vq_feat0"]], {1: st["mid1"], 0: st["mid0"]})
to = lambda t: t.to(dev)                                                    # noqa: E731
tc = decoder_train.DecoderTrainer(TorchLeaves(), sd)
tg = decoder_train.DecoderTrainer(encoder_train.CudaLeaves(make_dense("auto")), {k: to(v) for k, v in sd.items()})


def err(a, b):
    d = a - b.cpu()
    return float(d.norm()) / max(float(a.norm()), 1e-12), float(d.abs().max()) / max(float(a.abs().max()), 1e-12)


with torch.no_grad():
    rc = tc.forward(*args)
    rg = tg.forward(to(args[0]), [to(t) for t in args[1]], {k: to(v) for k, v in args[2].items()})
    for i, (a, b) in enumerate(zip(tc.vals, tg.vals)):
        e2, em = err(a, b)
        if e2 > 2e-5:
            print("fwd node %3d %-20s rel L2 %.2e max %.2e  mean %.3e" % (i, tuple(a.shape), e2, em, float(a.mean())))
    if "--force" in sys.argv:
        for ic, ig in zip(tc.offset_ids, tg.offset_ids):
            tg.vals[ig].copy_(tc.vals[ic].to(dev))
    seed = torch.randn(rc.shape, generator=torch.Generator().manual_seed(0))
    for graph, s in ((tc, seed), (tg, seed.to(dev))):
        gr, rec, T = {graph.out_id: s}, {}, graph.tape
        for kind, out, inputs, op in reversed(graph.nodes):
            gy = gr.pop(out, None)
            if gy is None:
                continue
            if kind == "add":
                vals = (gy, gy)
            elif kind == "fn":
                vals = op(gy)
            elif op[0] == "conv":
                vals = (T._conv_bwd(op[1], op[2], op[3], op[4], op[5], gy, op[6]),)
            elif op[0] == "gn":
                vals = (T._gn_bwd(op[1], op[2], op[3], op[4], gy),)
            else:
                vals = T._attn_bwd(op[1], op[2], op[3], gy)
            for i, v in zip(inputs, vals):
                if v is not None:
                    gr[i] = gr[i] + v if i in gr else v
                    rec[(out, i)] = v.detach().cpu().contiguous()
        graph.rec = rec
    last = 0.0
    for k in tc.rec:
        e2, em = err(tc.rec[k], tg.rec[k])
        node = tc.nodes[k[0] - 1]
        what = node[0] if node[0] != "op" else node[3][0] + " " + str(node[3][1])
        if e2 > 2 * last or e2 > 1e-3:
            print("bwd edge %-10s %-60s rel L2 %.2e max %.2e" % (k, what[:60], e2, em))
        last = max(last, e2)
    for k in sorted(tc.tape.grads):
        e2, em = err(tc.tape.grads[k], tg.tape.grads[k])
        if e2 > 5e-4:
            print("param %-60s rel L2 %.2e max %.2e" % (k, e2, em))
