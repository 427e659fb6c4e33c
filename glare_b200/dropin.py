"""Plug glare_b200 into the UNMODIFIED reference code base (LowLevelAI/GLARE ``code/``).

The reference selects its networks by name through two factories, ``models.networks.define_Flow`` (YAML
``network_G.which_model_G``) and ``models.networks.find_vqgan`` (``network_VQGAN.type``)
(code/models/networks.py:28-53, called from VQLLFLOWD_model.py:28-40), and calls the DCN operator through the
module-level name ``modulated_deform_conv`` (deformableDecoder_arch.py:151); stage 3 takes its two loss operators from
``models.modules.losses.PerceptualNetwork`` and ``models.modules.pytorch_msssim.msssim`` (VQLLFLOWD_model.py:16-17).  ``install()``
rebinds those names; entry points (``infer_unpaired.py``, ``infer_dataset_lol.py``), YAML files, checkpoints and the solver
classes stay untouched:

    import glare_b200.dropin as dropin
    dropin.install("/path/to/GLARE/code")          # before `from models import create_model`
    ... the reference's own main() ...

Only names are rebound; nothing is copied from or written to the reference tree.
"""
import importlib
import sys

from . import modules

_ARCH = {"VQLLFLOWDeformable": ("netG", modules.VQLLFLOWDeformable), "LLFlowVQGAN2": ("netG_stage2", modules.VQLLFLOWDeformable)}


def define_Flow(opt, step):
    """models/networks.py:28-36 replacement: same call signature, same returned interface."""
    opt_net = opt["network_G"]
    which = opt_net["which_model_G"]
    if which not in _ARCH:
        raise NotImplementedError("glare_b200 implements the shipped generators %s, not %r" % (sorted(_ARCH), which))
    key, cls = _ARCH[which]
    return cls(in_nc=opt_net["in_nc"], out_nc=opt_net["out_nc"], nf=opt_net["nf"], nb=opt_net["nb"], scale=opt["scale"],
               K=opt_net["flow"]["K"], opt=opt, step=step, which=key)


def find_vqgan(opt):
    """models/networks.py:38-53 replacement."""
    o = opt["network_VQGAN"]
    if o["type"] != "VQModel":
        raise NotImplementedError("glare_b200 implements network_VQGAN.type == 'VQModel', not %r" % (o["type"],))
    return modules.VQModel(resolution=o["resolution"], n_embed=o["n_embed"], z_channels=o["z_channels"], in_channels=o["in_channels"],
                           out_ch=o["out_ch"], ch=o["ch"], ch_mult=o["ch_mult"], num_res_blocks=o["num_res_blocks"],
                           attn_resolutions=o["attn_resolutions"])


def PerceptualNetwork():
    """losses.py:12-18 replacement: no arguments, pretrained torchvision VGG16 weights (or GLARE_VGG16_WEIGHTS), on the GPU"""
    from . import losses
    return losses.PerceptualNetwork(pretrained=True).to("cuda")


def install(reference_code_dir=None):
    """Rebind the reference's factory / operator names to the glare_b200 implementations.  Returns the patched modules."""
    if reference_code_dir and reference_code_dir not in sys.path:
        sys.path.insert(0, reference_code_dir)
    networks = importlib.import_module("models.networks")
    networks.define_Flow = define_Flow
    networks.find_vqgan = find_vqgan
    patched = [networks]
    for name in ("models.VQLLFLOWD_model", "models.LLFlow_model"):
        m = sys.modules.get(name)
        if m is not None and hasattr(m, "networks"):
            m.networks.define_Flow, m.networks.find_vqgan = define_Flow, find_vqgan
    try:
        dda = importlib.import_module("models.modules.deformableDecoder_arch")
        dda.modulated_deform_conv = modules.modulated_deform_conv
        dda.DCNv2Pack = modules.DCNv2Pack
        patched.append(dda)
    except Exception:       # the reference module imports its CUDA extension at import time; absence is not fatal here
        pass
    try:
        # stage 3 (VQLLFLOWD_model.py:16-17, 88-90): `from models.modules.losses import l1_loss, PerceptualNetwork` and
        # `from models.modules.pytorch_msssim import msssim` -- the two loss operators with a gradient for the reconstruction
        from . import losses
        ref_losses = importlib.import_module("models.modules.losses")
        ref_msssim = importlib.import_module("models.modules.pytorch_msssim")
        ref_losses.PerceptualNetwork = PerceptualNetwork
        ref_msssim.msssim = losses.msssim
        patched += [ref_losses, ref_msssim]
        m = sys.modules.get("models.VQLLFLOWD_model")
        if m is not None:
            m.PerceptualNetwork, m.msssim = PerceptualNetwork, losses.msssim
    except Exception:       # lpips / torchvision missing: the inference entry points do not need the losses
        pass
    return patched
