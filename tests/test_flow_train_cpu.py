"""CPU: host side of the stage-2 flow training step (glare_b200/flow_train.py: 28-step loops, offsets into the hoisted pre-activation
tensor, packed-parameter layout, assembly of the gradients into state-dict shapes) driven through the torch restatement of the kernel
contracts (tests/flow_train_emu.py), against the specification oracle/flow_backward.py -- which tests/test_oracle.py ties to autograd and
to the reference's own gradients."""
import torch
import torch.nn.functional as F

from flow_train_emu import TorchEmuKernels


def test_flow_training_step_host_logic_matches_specification(sd_g):
    from glare_b200 import flow, flow_train
    from oracle import flow_backward as FB
    gen = torch.Generator().manual_seed(21)
    B, h, w = 2, 5, 7
    gt = torch.randn((B, 3, h, w), generator=gen)
    ft = torch.sigmoid(torch.randn((B, 64, h, w), generator=gen))
    mean = torch.randn((B, 3, h, w), generator=gen) * 0.1
    plan = flow.FlowPlan(sd_g, torch.device("cpu"))
    conv = lambda x, wgt: F.conv2d(x, wgt, None, padding=1)               # noqa: E731
    with torch.no_grad():
        nll, z, g_gt, g_ft, g_mean, grads = flow_train.nll_forward_backward(plan, sd_g, gt, ft, mean, conv, kernels=TorchEmuKernels(sd_g))
        nll_s, z_s, g_gt_s, g_ft_s, g_mean_s, grads_s = FB.nll_forward_backward(sd_g, gt, ft, mean)

    def close(a, b, what):
        assert a.shape == b.shape, (what, a.shape, b.shape)
        sc = max(float(b.abs().max()), 1e-6)
        assert float((a - b).abs().max()) <= 2e-4 * sc + 1e-7, (what, float((a - b).abs().max()), sc)

    close(nll, nll_s, "nll"), close(z, z_s, "z"), close(g_gt, g_gt_s, "gt"), close(g_ft, g_ft_s, "ft"), close(g_mean, g_mean_s, "mean")
    assert sorted(grads) == sorted(grads_s)
    for k in grads_s:
        close(grads[k], grads_s[k], k)
        assert tuple(grads[k].shape) == tuple(sd_g[k].shape), k
