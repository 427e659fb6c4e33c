"""CPU: the drop-in module mirrors expose exactly the reference's state-dict keys/shapes (checkpoint contract,
SURVEY.md 5 'Checkpoint / resume'), accept the synthetic checkpoints with strict=True, and keep the reference's
constructor / forward signatures."""
import inspect
import os

import pytest
import torch

from glare_b200 import modules, synth


def test_netg_state_dict_contract(sd_g):
    net = modules.VQLLFLOWDeformable()
    shapes = synth.state_shapes("netG")
    assert sorted(net.state_dict().keys()) == sorted(shapes.keys())
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    net.load_state_dict(sd_g, strict=True)        # base_model.py:110-122 load_network(strict=True)
    k = "flowUpsamplerNet.layers.3.affine.fAffine.0.weight"
    assert torch.equal(net.state_dict()[k], sd_g[k])


def test_vqmodel_state_dict_contract(sd_v):
    net = modules.VQModel()
    shapes = synth.state_shapes("vqgan")
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    net.load_state_dict(sd_v, strict=True)


def test_stage2_generator_contract():
    net = modules.VQLLFLOWDeformable(which="netG_stage2")
    shapes = synth.state_shapes("netG_stage2")
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    net.load_state_dict(synth.synth_state_dict("netG_stage2", 0), strict=True)


def test_forward_signatures_match_reference():
    sig = inspect.signature(modules.VQLLFLOWDeformable.forward)
    assert list(sig.parameters)[1:8] == ["net_vq", "gt", "lr", "z", "eps_std", "reverse", "epses"]     # VQLLFLOWDeformable_arch.py:104
    sig = inspect.signature(modules.FlowUpsamplerNet.forward)
    assert list(sig.parameters)[1:] == ["gt", "rrdbResults", "z", "epses", "logdet", "reverse", "eps_std", "y_onehot"]  # FlowUpsamplerNet.py:208
    sig = inspect.signature(modules.modulated_deform_conv)
    assert list(sig.parameters) == ["input", "offset", "mask", "weight", "bias", "stride", "padding", "dilation", "groups",
                                    "deformable_groups"]                                                 # deform_conv.py:124-135
    sig = inspect.signature(modules.VectorQuantizer2.__init__)
    assert list(sig.parameters)[1:4] == ["n_e", "e_dim", "beta"]                                       # quantize.py:221


def test_dcn_pack_parameter_names():
    m = modules.DCNv2Pack(8, 8, 3, stride=1, padding=1, deformable_groups=4)
    assert sorted(m.state_dict()) == ["bias", "conv_offset.bias", "conv_offset.weight", "weight"]
    assert m.conv_offset.weight.shape == (108, 8, 3, 3) and float(m.conv_offset.weight.abs().max()) == 0.0   # deform_conv.py:367-371


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="reference tree not mounted")
def test_key_list_matches_live_reference():
    """state_shapes.json (what the mirrors are built from) against the reference constructors themselves"""
    from oracle import ref_shims
    netG, net_hq, _ = ref_shims.build_reference("LOL.yml", seed=0)
    assert {k: list(v.shape) for k, v in netG.state_dict().items()} == synth.state_shapes("netG")
    assert {k: list(v.shape) for k, v in net_hq.state_dict().items()} == synth.state_shapes("vqgan")


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="reference tree not mounted")
def test_dropin_rebinds_reference_factories(sd_g, sd_v):
    """install() swaps the reference's name-based factories; the reference's own option parser feeds them."""
    from oracle import ref_shims
    from glare_b200 import dropin
    ref_shims.install()
    import models.networks as networks
    saved = (networks.define_Flow, networks.find_vqgan)
    import models.modules.deformableDecoder_arch as dda
    saved_dcn = (dda.modulated_deform_conv, dda.DCNv2Pack)
    import models.modules.losses as ref_losses
    import models.modules.pytorch_msssim as ref_msssim
    saved_losses = (ref_losses.PerceptualNetwork, ref_msssim.msssim)
    try:
        dropin.install()
        opt = ref_shims.parse_opt("LOL.yml")
        netG = networks.define_Flow(opt, 0)
        net_hq = networks.find_vqgan(opt)
        assert isinstance(netG, modules.VQLLFLOWDeformable) and isinstance(net_hq, modules.VQModel)
        netG.load_state_dict(sd_g, strict=True)
        net_hq.load_state_dict(sd_v, strict=True)
        assert dda.modulated_deform_conv is modules.modulated_deform_conv
        assert not any(p.requires_grad for p in netG.RRDB.parameters()) and not any(p.requires_grad for p in netG.flowUpsamplerNet.parameters())
        assert all(p.requires_grad for p in netG.deformable_decoder.parameters())          # VQLLFLOWDeformable_arch.py:49-52
        from glare_b200 import losses
        assert ref_msssim.msssim is losses.msssim and ref_losses.PerceptualNetwork is dropin.PerceptualNetwork
    finally:
        networks.define_Flow, networks.find_vqgan = saved
        dda.modulated_deform_conv, dda.DCNv2Pack = saved_dcn
        ref_losses.PerceptualNetwork, ref_msssim.msssim = saved_losses


def test_deform_conv_ext_shim_exports_the_reference_plugin_surface():
    """the five functions of the reference's pybind11 module (ops/dcn/src/deform_conv_ext.cpp:150-164) with the reference's argument counts"""
    import inspect as ins
    from glare_b200 import deform_conv_ext as ext
    want = {"deform_conv_forward", "deform_conv_backward_input", "deform_conv_backward_parameters", "modulated_deform_conv_forward",
            "modulated_deform_conv_backward"}
    assert want <= set(dir(ext))
    assert len(ins.signature(ext.modulated_deform_conv_forward).parameters) == 19       # deform_conv_ext.cpp:125-134
    assert len(ins.signature(ext.modulated_deform_conv_backward).parameters) == 24      # :136-147
    x = torch.zeros((1, 8, 4, 4))
    with pytest.raises(NotImplementedError):                                             # CPU tensors: as the reference
        ext.modulated_deform_conv_forward(x, torch.zeros((8, 8, 3, 3)), torch.zeros(8), x.new_empty(0), torch.zeros((1, 72, 4, 4)),
                                          torch.zeros((1, 36, 4, 4)), torch.zeros((1, 8, 4, 4)), x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, 4, True)


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="reference tree not mounted")
def test_reference_function_binds_the_shim_without_rebinding():
    """registering the shim under the name the reference imports (deform_conv.py:23-26 `from . import deform_conv_ext`) makes the
    reference's OWN ModulatedDeformConvFunction call into it: checked on CPU up to the shim's NotImplementedError for CPU tensors"""
    import importlib
    import sys
    from oracle import ref_shims
    from glare_b200 import deform_conv_ext as ext
    ref_shims.install()
    name = "models.modules.ops.dcn.deform_conv_ext"
    saved, saved_mod = sys.modules.get(name), sys.modules.pop("models.modules.ops.dcn.deform_conv", None)
    try:
        sys.modules[name] = ext
        dc = importlib.import_module("models.modules.ops.dcn.deform_conv")
        assert dc.deform_conv_ext is ext
        x = torch.zeros((1, 8, 4, 4))
        with pytest.raises(NotImplementedError):
            dc.modulated_deform_conv(x, torch.zeros((1, 72, 4, 4)), torch.zeros((1, 36, 4, 4)), torch.zeros((8, 8, 3, 3)), torch.zeros(8), 1, 1, 1, 1, 4)
    finally:
        sys.modules.pop(name, None)
        sys.modules.pop("models.modules.ops.dcn.deform_conv", None)
        if saved is not None:
            sys.modules[name] = saved
        if saved_mod is not None:
            sys.modules["models.modules.ops.dcn.deform_conv"] = saved_mod
