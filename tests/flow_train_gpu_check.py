"""Child process of tests/test_zz_flow_train_gpu.py: the stage-2 flow training kernels (csrc/flow_bwd.cu) on cuda:0, kernel by kernel against
the torch restatement of their contracts (tests/flow_train_emu.py) and end to end against the CPU specification (oracle/flow_backward.py).
Prints one line per check and exits 0 only if every check holds.  Runs in its own process so that a faulting kernel cannot take the
CUDA context of the main test run with it."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from flow_train_emu import TorchEmuKernels  # noqa: E402
from glare_b200 import flow, flow_train, synth  # noqa: E402
from glare_b200.flow import COUPLING_STEPS  # noqa: E402
from oracle import flow_backward as FB  # noqa: E402

OK = True


def check(name, a, b, tol=2e-4):
    global OK
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    sc = max(float(b.abs().max()), 1e-6)
    err = float((a - b).abs().max()) if a.shape == b.shape else float("inf")
    good = err <= tol * sc + 1e-7 and bool(torch.isfinite(a).all())
    OK = OK and good
    print("%-4s %-58s err %.3g (scale %.3g)" % ("ok" if good else "FAIL", name, err, sc), flush=True)


def run_checks(dev, K, E, conv, tol_step=1.0, shape=(2, 9, 11)):
    """K: kernels under test, E: torch restatement of their contracts (same device); conv(x, w): dense 3x3 'same' conv"""
    sd = synth.synth_state_dict("netG", 0)
    plan = flow.FlowPlan(sd, dev)
    gen = torch.Generator().manual_seed(33)
    B, h, w = shape
    P = B * h * w
    gt = torch.randn((B, 3, h, w), generator=gen)
    ft = torch.sigmoid(torch.randn((B, 64, h, w), generator=gen))
    mean = torch.randn((B, 3, h, w), generator=gen) * 0.1
    rnd = lambda *s: torch.randn(s, generator=gen).to(dev)             # noqa: E731
    new = lambda *s: torch.full(s, float("nan"), device=dev)           # noqa: E731

    # ---- kernel by kernel (same inputs through the CUDA kernel and through the torch restatement of its contract)
    ci = 3
    netA, netF, pw = plan.nets_a[ci], plan.nets_f[ci], plan.pw_fwd[COUPLING_STEPS[ci]]
    pre = rnd(B, h, w, 256)
    zin = rnd(B, 3, h, w)
    for tag, net, z1 in (("NN_A", netA, rnd(P, 4)), ("NN_F", netF, None)):
        outs = [[new(P, 64), new(P, 64), new(P, 8)] for _ in range(2)]
        for kern, o in ((K, outs[0]), (E, outs[1])):
            kern.net_fwd(pre, 128, 256, z1, 4, net, B, h, w, *o)
        for nm, a, b in zip(("h1", "h2", "hout"), *outs):
            check("net_fwd %s %s" % (tag, nm), a, b)
        h1, h2, hout = outs[1]
        g_h = rnd(P, 8)
        g_h[:, (4 if z1 is not None else 6):] = 0
        res = []
        for kern in (K, E):
            bufs = [new(P, 8), new(P, 64), new(P, 64), new(P, 64), new(P, 64)]
            g_pre, g_z1 = torch.zeros((B, h, w, 256), device=dev), (new(P, 1) if z1 is not None else None)
            kern.net_bwd(g_h, h1, h2, net, B, h, w, *bufs, g_pre, 64, 256, g_z1)
            res.append(bufs + [g_pre] + ([g_z1] if g_z1 is not None else []))
        for nm, a, b in zip(("g_a3", "g_n2", "g_a2", "g_n1", "g_a1", "g_pre", "g_z1"), *res):
            check("net_bwd %s %s" % (tag, nm), a, b)
    hF = rnd(P, 8)
    for hf, tag in ((hF, "coupling"), (None, "noCoupling")):
        outs = [[new(P, 4), new(P, 4), new(P, 4)] for _ in range(2)]
        for kern, o in ((K, outs[0]), (E, outs[1])):
            kern.point_fwd(zin, pw, hf, B, h, w, *o)
        for nm, a, b in zip("tuv", *outs):
            check("point_fwd %s %s" % (tag, nm), a, b)
    t, u, v = outs[1]
    hA, g_out, g_z1 = rnd(P, 8), rnd(B, 3, h, w), rnd(P, 1)
    for which, g_in, gz, x, hr in ((0, g_out, None, v, hA), (1, rnd(P, 4), g_z1, u, hF)):
        outs = [[new(P, 8), new(P, 4)] for _ in range(2)]
        for kern, o in ((K, outs[0]), (E, outs[1])):
            kern.coupling_bwd(which, g_in, gz, x, hr, -0.0123, B, h, w, *o)
        check("coupling_bwd %d g_h" % which, outs[0][0], outs[1][0])
        check("coupling_bwd %d g_x" % which, outs[0][1][:, :3], outs[1][0 + 1][:, :3])
    g_u = rnd(P, 4)
    outs = [[new(B, 3, h, w), torch.zeros(16, device=dev)] for _ in range(2)]
    for kern, o in ((K, outs[0]), (E, outs[1])):
        kern.point_bwd(g_u, t, pw, B, h, w, *o)
    check("point_bwd g_z", outs[0][0], outs[1][0])
    check("point_bwd sums", outs[0][1][:15], outs[1][1][:15])
    for Cx, ldx in ((64, 64), (1, 4)):
        x = rnd(P, ldx)
        outs = [new(P, 9 * Cx) for _ in range(2)]
        for kern, o in ((K, outs[0]), (E, outs[1])):
            kern.im2col3x3(x, ldx, Cx, B, h, w, o)
        check("im2col3x3 C=%d" % Cx, *outs)
    a, b = rnd(P, 64), rnd(P, 64)
    for Cx, bb in ((64, b), (64, None), (8, None)):
        outs = [torch.zeros(64, device=dev) for _ in range(2)]
        for kern, o in ((K, outs[0]), (E, outs[1])):
            kern.colsum(a, 64 if Cx == 64 else 8, bb, 64, Cx, P if Cx == 64 else P * 8, o)
        check("colsum C=%d%s" % (Cx, " dot" if bb is not None else ""), *outs)
    outs = [torch.zeros((9, 64), device=dev) for _ in range(2)]
    c9 = rnd(P, 9)
    for kern, o in ((K, outs[0]), (E, outs[1])):
        kern.gemm_tn(c9, 9, a, 64, P, o)
    check("gemm_tn 9 x 64", *outs)

    # ---- the whole training step: kernels + the dense conv path against the CPU specification
    with torch.no_grad():
        nll, z, g_gt, g_ft, g_mean, grads = flow_train.nll_forward_backward(plan, sd, gt.to(dev), ft.to(dev), mean.to(dev), conv, kernels=K)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        nll_s, z_s, g_gt_s, g_ft_s, g_mean_s, grads_s = FB.nll_forward_backward(sd, gt, ft, mean)
    check("step nll", nll, nll_s, 1e-4 * tol_step)
    check("step z", z, z_s, 1e-3 * tol_step)
    check("step dL/dgt", g_gt, g_gt_s, 2e-3 * tol_step)
    check("step dL/dft", g_ft, g_ft_s, 2e-3 * tol_step)
    check("step dL/dmean", g_mean, g_mean_s, 2e-3 * tol_step)
    # a pre-activation within rounding of zero can land on the other side of the ReLU in a different evaluation order (the hoisted conv splits
    # the first layer's sum): such a flip changes one output channel of one net's gradients by ~1e-3 of their scale.  Allow a few such tensors.
    rel = sorted(((float((grads[k].cpu() - grads_s[k]).abs().max()) / max(float(grads_s[k].abs().max()), 1e-6), k) for k in grads_s), reverse=True)
    outliers = [r for r in rel if r[0] >= 5e-4]
    good = sorted(grads) == sorted(grads_s) and len(outliers) <= 6 and rel[0][0] < 5e-2
    print("%-4s step parameter gradients: %d tensors, %d above 5e-4 of their scale, worst %.3g at %s" %
          ("ok" if good else "FAIL", len(grads_s), len(outliers), rel[0][0], rel[0][1]))
    return OK and good


def main():
    dev = torch.device("cuda:0")
    # the torch restatement of the kernel contracts runs on the GPU through cuDNN / cuBLAS: TF32 must be off there (cuDNN convs allow it by
    # default), or the REFERENCE side of each check carries 3e-4 of error (round 2's first hardware run: 13 such false failures)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd_d = {k: v.to(dev) for k, v in synth.synth_state_dict("netG", 0).items()}
    from glare_b200.dense import make_dense
    dense = make_dense("auto")
    conv = lambda x, wgt: dense.conv2d(x, wgt).float()                  # noqa: E731
    ok = run_checks(dev, flow_train.CudaKernels(), TorchEmuKernels(sd_d), conv)
    ok = encoder_step_check(dev, dense, conv) and ok
    sys.exit(0 if ok else 1)


def encoder_step_check(dev, dense, conv):
    """the whole stage-2 step (tape over ConEncoder1 + flow objective) on the GPU kernels against the same host logic on torch primitives (CPU),
    which tests/test_encoder_train_cpu.py ties to autograd and to the reference's gradients"""
    import torch.nn.functional as F
    from encoder_train_emu import TorchLeaves
    from glare_b200 import encoder_train
    sd = synth.synth_state_dict("netG_stage2", 0)
    gen = torch.Generator().manual_seed(5)
    gt = torch.randn((2, 3, 8, 8), generator=gen)
    lr = synth.preprocess(torch.rand((2, 3, 32, 32), generator=gen))
    with torch.no_grad():
        nll_c, grads_c = encoder_train.stage2_step(sd, flow.FlowPlan(sd, torch.device("cpu")), lr, gt, TorchLeaves(),
                                                   lambda x, wgt: F.conv2d(x, wgt, None, padding=1), flow_kernels=TorchEmuKernels(sd))
        sd_d = {k: v.to(dev) for k, v in sd.items()}
        nll_g, grads_g = encoder_train.stage2_step(sd_d, flow.FlowPlan(sd, dev), lr.to(dev), gt.to(dev), encoder_train.CudaLeaves(dense), conv)
        torch.cuda.synchronize()
    check("stage-2 step nll", nll_g, nll_c, 1e-4)
    # opt-in tensor-core weight gradient (chunked reduction as the batch of one GEMM with per-sample weights) against fp64
    Lg = encoder_train.CudaLeaves(dense)
    for (Pp, M, N) in ((5000, 1152, 128), (777, 128, 64)):
        a, b = torch.randn((Pp, M), generator=gen).to(dev), torch.randn((Pp, N), generator=gen).to(dev)
        check("gemm_tn_tc %dx%dx%d" % (Pp, M, N), Lg.gemm_tn_tc(a, b, chunk=1024), (a.double().t() @ b.double()).float(), 1e-4)
    # left operand given as its transpose (attention backward: P^T dO, dS^T Q), K not a multiple of the K granule
    for (Kk, hh, ww, Nn) in ((200, 8, 8, 40), (4096, 16, 16, 512)):
        at, bq = torch.randn((Kk, hh * ww), generator=gen), torch.randn((Nn, Kk), generator=gen)
        check("gemm_nt_ta K=%d R=%d N=%d" % (Kk, hh * ww, Nn), Lg.gemm_nt_ta(at.to(dev), bq.to(dev), (hh, ww)), (at.double().t() @ bq.double().t()).float(), 1e-4)
    a27, b128 = torch.randn((5003, 27), generator=gen), torch.randn((5003, 128), generator=gen)
    check("gemm_tn skinny 27 x 128 (conv_in)", Lg.gemm_tn(a27.to(dev), b128.to(dev)), (a27.double().t() @ b128.double()).float(), 2e-5)
    for (Pp, C) in ((204800, 128), (777, 512), (51, 4)):
        xs = torch.randn((Pp, C), generator=gen)
        check("colsum %dx%d" % (Pp, C), Lg.colsum(xs.to(dev)), xs.double().sum(dim=0).float(), 2e-5)
    # a 3-output-channel conv (residual_conv / color_conv): dY zero-padded to 32 columns inside wgrad_conv
    xq, gq = torch.randn((2, 20, 24, 128), generator=gen), torch.randn((2, 20, 24, 3), generator=gen)
    want = (TorchLeaves().im2col(xq, 3, 1, 1, 20, 24).double().t() @ gq.reshape(-1, 3).double()).float()
    check("wgrad_conv Co=3", Lg.wgrad_conv(xq.to(dev), gq.to(dev), 3, 1, 1), want, 1e-4)
    # im2col-free tensor-core weight gradient (csrc/train_wgrad.cu: transposed operands written directly) against fp64 over the torch im2col
    T = TorchLeaves()
    for (Bq, Hh, Ww, Ci, Co, kk, stride, pad) in ((2, 13, 21, 128, 128, 3, 1, 1), (1, 22, 18, 64, 128, 3, 2, 0), (3, 9, 11, 256, 64, 1, 1, 0),
                                                  (1, 40, 40, 32, 32, 3, 1, 1)):
        Ho, Wo = (Hh, Ww) if stride == 1 else ((Hh + 1 - 3) // 2 + 1, (Ww + 1 - 3) // 2 + 1)
        xq = torch.randn((Bq, Hh, Ww, Ci), generator=gen)
        gq = torch.randn((Bq, Ho, Wo, Co), generator=gen)
        want = (T.im2col(xq, kk, stride, pad, Ho, Wo).double().t() @ gq.reshape(-1, Co).double()).float()
        for chunk in (8192, 96):                        # one chunk / many ragged chunks
            got = Lg.wgrad_conv(xq.to(dev), gq.to(dev), kk, stride, pad, chunk=chunk)
            check("wgrad_conv %s chunk %d" % ((Bq, Hh, Ww, Ci, Co, kk, stride), chunk), got, want, 1e-4)
    rel = sorted(((float((grads_g[k].cpu() - grads_c[k]).abs().max()) / max(float(grads_c[k].abs().max()), 1e-6), k) for k in grads_c), reverse=True)
    # first hardware run of a 40-layer backward through fp32-grade (bf16x3) tensor-core convs against fp32 CPU arithmetic: report the error
    # distribution, require it to be small for all but a few tensors (isolated ReLU sign flips of the flow nets, see run_checks)
    import statistics
    outliers = [r for r in rel if r[0] >= 1e-2]
    good = sorted(grads_g) == sorted(grads_c) and len(outliers) <= max(8, len(rel) // 20) and rel[0][0] < 0.2
    print("%-4s stage-2 step gradients: %d tensors, median relative error %.3g, %d above 1e-2 of their scale" %
          ("ok" if good else "FAIL", len(grads_c), statistics.median(r[0] for r in rel), len(outliers)))
    for e, k in rel[:10]:
        print("       %.3g  %s" % (e, k))
    # the bf16 training configuration (BASELINE config 4 names bf16): single-piece bf16 operands in every convolution, attention GEMM and
    # weight gradient (csrc/train_wgrad.cu's mode-0 operand), fp32 accumulation and fp32 memory-bound kernels.  Against the fp32-grade step
    # on the same device: the objective within 1e-3, the gradients within bf16's 2^-9 per operand (relative L2 per tensor)
    from glare_b200 import flow_train
    from glare_b200.dense import make_dense
    d16 = make_dense("tc-bf16")
    with torch.no_grad():
        nll_b, grads_b = encoder_train.stage2_step(sd_d, flow.FlowPlan(sd, dev), lr.to(dev), gt.to(dev), encoder_train.CudaLeaves(d16),
                                                   lambda x, wgt: d16.conv2d(x, wgt).float(), flow_kernels=flow_train.CudaKernels(mode=0))
        torch.cuda.synchronize()
    check("stage-2 step nll, bf16 operands", nll_b, nll_g, 1e-3)
    gmax = max(float(v.abs().max()) for v in grads_g.values())
    rel16 = sorted(float((grads_b[k] - v).norm()) / max(float(v.norm()), 1e-4 * gmax * v.numel() ** 0.5) for k, v in grads_g.items())
    good16 = sorted(grads_b) == sorted(grads_g) and statistics.median(rel16) < 1e-2 and rel16[len(rel16) * 9 // 10] < 5e-2
    print("%-4s stage-2 step gradients, bf16 operands vs fp32-grade: relative L2 per tensor median %.3g, 90th percentile %.3g, max %.3g" %
          ("ok" if good16 else "FAIL", statistics.median(rel16), rel16[len(rel16) * 9 // 10], rel16[-1]))
    return OK and good and good16


if __name__ == "__main__":
    main()
