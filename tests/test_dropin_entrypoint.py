"""CPU (authoring container only: needs /root/reference): the UNMODIFIED reference entry point `infer_unpaired.main()`
(code/infer_unpaired.py:91-136; BASELINE.json configs[0]) driven through `glare_b200.dropin.install()` on synthetic checkpoints.

What runs is the reference's own code -- option parser, `create_model`, `VQLLFLOWDModel.__init__` (torch.load of both checkpoints),
`load_network(strict=True)`, `auto_padding`, `get_sr` -> `get_sr_with_z` -> `get_z` (reads netG.flowUpsamplerNet.scaleH / scaleW), the
generator call `netG(net_vq=net_hq, lr=..., z=..., eps_std=..., reverse=True, ...)`, crop, `imwrite` -- with the generator / VQGAN /
DCN names rebound to the glare_b200 mirrors.  The mirrors' engine is CUDA-only and this container has no GPU, so the test swaps the
ENGINE (and only the engine) for a stand-in that evaluates the CPU oracle: everything up to and after the kernel boundary is the product's
and the reference's real code, and the written PNG is compared with the oracle run through the reference's pre/post-processing steps.
Without the stand-in the same call must fail loudly (no CPU fallback), which is asserted too."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="reference tree not mounted")


class _OracleEngine:
    """TEST-ONLY stand-in for glare_b200.engine.GlareEngine (same constructor / infer signature), evaluating oracle/glare_oracle.py"""

    def __init__(self, sd_g, sd_vq, device="cpu", dense=None, per_sample_ratio=True, flow=True, decoders=True):
        self.sd_g = {k: v.detach().float().cpu() for k, v in sd_g.items()}
        self.sd_v = {k: v.detach().float().cpu() for k, v in sd_vq.items()}
        self.device = torch.device("cpu")

    def infer(self, lr, stages=None, graph=False):
        from oracle import glare_oracle as O
        st = {}
        out = O.glare_infer(self.sd_g, self.sd_v, lr.float().cpu(), stages=st)
        if stages is not None:
            stages.update(st)
        return out


def _write_case(tmp, sd_g, sd_v):
    import cv2
    import yaml
    from glare_b200 import synth
    os.makedirs(os.path.join(tmp, "imgs"))
    os.makedirs(os.path.join(tmp, "code"))
    torch.save(sd_g, os.path.join(tmp, "net_G.pth"))          # plain state_dicts (base_model.py:101-108 format)
    torch.save(sd_v, os.path.join(tmp, "vqgan.pkl"))
    lq, _ = synth.synth_images(1, 40, 56, seed=4)
    rgb = (lq[0].permute(1, 2, 0) * 255.0).round().to(torch.uint8).numpy()
    cv2.imwrite(os.path.join(tmp, "imgs", "a.png"), rgb[:, :, ::-1])
    with open("/root/reference/code/confs/LOL.yml") as f:
        y = yaml.safe_load(f)
    y["dataroot_unpaired"] = os.path.join(tmp, "imgs")        # commented out in LOL.yml:62; no new keys otherwise
    y["model_path"] = os.path.join(tmp, "net_G.pth")
    y["path"]["pretrain_model_G"] = os.path.join(tmp, "net_G.pth")
    y["path"]["pretrained_vqgan"] = os.path.join(tmp, "vqgan.pkl")
    conf = os.path.join(tmp, "LOL.yml")
    with open(conf, "w") as f:
        yaml.safe_dump(y, f)
    return conf, rgb


def test_infer_unpaired_main_through_dropin(tmp_path, sd_g, sd_v, monkeypatch):
    import cv2
    from glare_b200 import dropin, modules, synth
    from oracle import glare_oracle as O
    from oracle import ref_shims
    tmp = str(tmp_path)
    conf, rgb = _write_case(tmp, sd_g, sd_v)
    ref_shims.install()
    import models.networks as networks
    import models.modules.deformableDecoder_arch as dda
    saved = (networks.define_Flow, networks.find_vqgan, dda.modulated_deform_conv, dda.DCNv2Pack)
    import models.modules.losses as ref_losses
    import models.modules.pytorch_msssim as ref_msssim
    saved_losses = (ref_losses.PerceptualNetwork, ref_msssim.msssim)
    cwd = os.getcwd()
    try:
        dropin.install()
        entry = importlib.import_module("infer_unpaired")
        # the script writes next to its own __file__ (infer_unpaired.py:106-107): point that away from the read-only reference tree
        monkeypatch.setattr(entry, "__file__", os.path.join(tmp, "code", "infer_unpaired.py"))
        monkeypatch.setattr(sys, "argv", ["infer_unpaired.py", "--opt", conf, "-n", "t"])
        # 1. the product as shipped: CUDA-only, fails loudly on this GPU-less host (no CPU fallback behind the drop-in)
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match="CUDA"):
                entry.main()
        # 2. engine swapped for the oracle stand-in: the entry point runs end to end
        monkeypatch.setattr(modules, "GlareEngine", _OracleEngine)
        entry.main()
    finally:
        os.chdir(cwd)
        networks.define_Flow, networks.find_vqgan, dda.modulated_deform_conv, dda.DCNv2Pack = saved
        ref_losses.PerceptualNetwork, ref_msssim.msssim = saved_losses
        vm = sys.modules.get("models.VQLLFLOWD_model")
        if vm is not None:
            vm.PerceptualNetwork, vm.msssim = saved_losses
    out_path = os.path.join(tmp, "results-unpair", "LOL", "t", "a.png")
    assert os.path.exists(out_path)
    got = cv2.imread(out_path)[:, :, ::-1]
    assert got.shape == rgb.shape
    # the same image through the reference's steps restated in synth + the oracle
    x = torch.from_numpy(rgb.copy()).permute(2, 0, 1)[None].float() / 255.0
    xp, (h1, h2, w1, w2) = synth.auto_padding(x)
    ref = O.glare_infer(sd_g, sd_v, synth.preprocess(xp))
    ref = (ref[:, :, h1:ref.shape[2] - h2, w1:ref.shape[3] - w2].clamp(0, 1) * 255.0).to(torch.uint8)[0].permute(1, 2, 0).numpy()
    assert int(np.abs(got.astype(np.int32) - ref.astype(np.int32)).max()) <= 1


def test_get_z_attributes_present():
    """VQLLFLOWD_model.py:307-321 reads these from the generator (AttributeError on the first image)"""
    from glare_b200 import modules
    from oracle import ref_shims
    opt = ref_shims.parse_opt("LOL.yml")
    net = modules.VQLLFLOWDeformable(opt=opt)
    f = net.flowUpsamplerNet
    assert (f.H, f.W, f.C) == (80, 80, 3) and f.scaleH == pytest.approx(256 / 80) and f.scaleW == pytest.approx(256 / 80)
    assert net.quant == 32 and net.opt is opt
