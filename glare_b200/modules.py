"""Drop-in nn.Module mirrors of the reference classes on the hot path: same constructor/forward signatures,
same parameter names and shapes (so ``net_G.pth`` / ``vqgan.pkl`` load with ``strict=True``,
code/models/base_model.py:110-122), kernels from libglare_b200.so underneath.

    VectorQuantizer2        code/models/modules/quantize.py:213-312
    FlowUpsamplerNet        code/models/modules/FlowUpsamplerNet.py:18-326
    VQLLFLOWDeformable      code/models/modules/VQLLFLOWDeformable_arch.py:18-250   (netG of inference / stage 3)
    VQModel                 code/models/modules/VQModel_arch.py:14-110              (net_hq)
    modulated_deform_conv / ModulatedDeformConvPack / DCNv2Pack
                            code/models/modules/ops/dcn/deform_conv.py:121-188,300-377, deformableDecoder_arch.py:122-152

The sub-trees whose arithmetic lives in the engine (encoders, decoders) are ``ParamModule``s: parameter
containers built from the reference's own state-dict key list (glare_b200/data/state_shapes.json), not
re-implementations of the reference forward code.  In training mode the generator's two training calls run on the library's kernels too:
``reverse=False`` (stage 2: the flow objective, encoder_train.py) and ``reverse=True, reverse_with_grad=True`` (stage 3: the deformable decoder,
decoder_train.py); the reference's solvers (optimizers, schedulers, logging) stay as they are.
"""
import math

import torch
import torch.nn as nn

from . import flow as flowmod
from . import ops, synth
from .engine import GlareEngine


class ParamModule(nn.Module):
    """Parameter container whose ``state_dict()`` keys equal the given key list (nested by '.')."""

    def __init__(self, shapes):
        super().__init__()
        groups = {}
        for key, shape in shapes.items():
            head, _, rest = key.partition(".")
            if rest:
                groups.setdefault(head, {})[rest] = shape
            else:
                self.register_parameter(head, nn.Parameter(torch.zeros(tuple(shape))))
        for head, sub in groups.items():
            self.add_module(head, ParamModule(sub))


def _sub(shapes, prefix):
    return {k[len(prefix):]: v for k, v in shapes.items() if k.startswith(prefix)}


def _fingerprint(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


# ----------------------------------------------------------------------------------------------- VQ
class VectorQuantizer2(nn.Module):
    """quantize.py:213-312.  forward(z) -> (z_q, loss, (None, None, indices)); indices bit-exact with the reference's
    CPU fp32 evaluation, z_q = z + (e - z) (straight-through, :298)."""

    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=False):
        super().__init__()
        if remap is not None:
            raise NotImplementedError("remap is not used by any GLARE configuration (LOL.yml network_VQGAN)")
        self.n_e, self.e_dim, self.beta, self.legacy = n_e, e_dim, beta, legacy
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)
        self.sane_index_shape = sane_index_shape
        self._packed, self._packed_key = None, None

    def _codebook(self):
        key = (self.embedding.weight.data_ptr(), self.embedding.weight._version)
        if key != self._packed_key:
            self._packed, self._packed_key = ops.vq_pack_codebook(self.embedding.weight.detach()), key
        return self._packed

    def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
        assert temp is None or temp == 1.0, "Only for interface compatible with Gumbel"
        assert rescale_logits is False and return_logits is False, "Only for interface compatible with Gumbel"
        idx, z_q = ops.vq_lookup(z.detach(), self._codebook())
        zd = z.detach()
        # quantize.py:290-295 (commitment loss; legacy=False by default in VQModel_arch.py)
        if not self.legacy:
            loss = self.beta * torch.mean((z_q - zd) ** 2) + torch.mean((z_q - zd) ** 2)
        else:
            loss = torch.mean((z_q - zd) ** 2) + self.beta * torch.mean((z_q - zd) ** 2)
        if self.sane_index_shape:
            idx = idx.reshape(z_q.shape[0], z_q.shape[2], z_q.shape[3])
        return z_q, loss, (None, None, idx)

    def get_codebook_entry(self, indices, shape):
        z_q = self.embedding(indices)
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
        return z_q


# ----------------------------------------------------------------------------------------------- DCN
def modulated_deform_conv(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                          deformable_groups=1):
    """ops/dcn/deform_conv.py:188 ``modulated_deform_conv = ModulatedDeformConvFunction.apply`` (forward and backward)."""
    from .dcn_backward import ModulatedDeformConvFunction
    return ModulatedDeformConvFunction.apply(input, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups)


class ModulatedDeformConvPack(nn.Module):
    """ops/dcn/deform_conv.py:300-377: owns weight/bias and the zero-initialised ``conv_offset`` (:367-371)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        k = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, k
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.deformable_groups, self.with_bias = deformable_groups, bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *k))
        self.bias = nn.Parameter(torch.Tensor(out_channels)) if bias else None
        n = in_channels * k[0] * k[1]
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()
        self.conv_offset = nn.Conv2d(in_channels, deformable_groups * 3 * k[0] * k[1], kernel_size=k, stride=stride,
                                     padding=padding, dilation=dilation, bias=True)
        self.conv_offset.weight.data.zero_()
        self.conv_offset.bias.data.zero_()

    def forward(self, x):
        out = self.conv_offset(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        return modulated_deform_conv(x, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias, self.stride,
                                     self.padding, self.dilation, self.groups, self.deformable_groups)


class DCNv2Pack(ModulatedDeformConvPack):
    """deformableDecoder_arch.py:122-152: offsets and masks come from a second feature map."""

    def forward(self, x, feat):
        out = self.conv_offset(feat).to(torch.float32)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        return modulated_deform_conv(x, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias, self.stride,
                                     self.padding, self.dilation, self.groups, self.deformable_groups)


# ----------------------------------------------------------------------------------------------- flow
class FlowUpsamplerNet(ParamModule):
    """FlowUpsamplerNet.py:18-326 for the shipped configuration (LOL.yml: K=12, L=2, 3-channel latent, no split).
    forward(gt=None, rrdbResults=None, z=None, epses=None, logdet=0., reverse=False, eps_std=None, y_onehot=None)
    -> (tensor, logdet)."""

    def __init__(self, image_shape=None, hidden_channels=64, K=12, L=None, actnorm_scale=1.0, flow_permutation=None,
                 flow_coupling="affine", LU_decomposed=False, opt=None, shapes=None):
        super().__init__(shapes if shapes is not None else _sub(synth.state_shapes("netG"), "flowUpsamplerNet."))
        # attributes the reference solver reads (VQLLFLOWD_model.py:307-321 get_z): FlowUpsamplerNet.py:38,116-119 with the generator's
        # image_shape (80, 80, 3) (VQLLFLOWDeformable_arch.py:45) and no squeeze level in the shipped configuration
        self.H, self.W, self.C = (tuple(image_shape) if image_shape is not None else (80, 80, 3))
        gt_size = 256
        try:
            gt_size = opt["datasets"]["train"]["GT_size"] or gt_size
        except (TypeError, KeyError):
            pass
        self.scaleH, self.scaleW = gt_size / self.H, gt_size / self.W
        self.opt = opt
        self._plan, self._plan_key = None, None
        self.dense = None          # dense conv backend for the hoisted first layers; set by the owning generator

    def plan(self):
        key = _fingerprint(self)
        if key != self._plan_key:
            sd = {"flowUpsamplerNet." + k: v for k, v in self.state_dict().items()}
            self._plan, self._plan_key = flowmod.FlowPlan(sd, next(self.parameters()).device), key
        return self._plan

    def _conv(self, x, w):
        if self.dense is None:
            from .dense import make_dense
            self.dense = make_dense("auto")
        return self.dense.conv2d(x, w).float()

    def forward(self, gt=None, rrdbResults=None, z=None, epses=None, logdet=0., reverse=False, eps_std=None, y_onehot=None):
        ft = rrdbResults if isinstance(rrdbResults, torch.Tensor) else rrdbResults["cond_feat"]   # FlowUpsamplerNet.py:68-71
        src = z if reverse else gt
        assert src is not None
        ld = logdet if isinstance(logdet, torch.Tensor) else torch.full((src.shape[0],), float(logdet), device=src.device)
        ld = ld.float().clone()
        if reverse:
            return flowmod.decode(self.plan(), src, ft, self._conv, logdet=ld)
        return flowmod.encode(self.plan(), src, ft, self._conv, logdet=ld)


# ----------------------------------------------------------------------------------------------- generator + VQGAN
class VQModel(nn.Module):
    """VQModel_arch.py:14-110 (net_hq).  encode(x) -> (h, None); decode(h) -> (None, emb_loss, [feat@H/2, feat@H]);
    the RGB reconstruction the reference computes and discards (VQLLFLOWDeformable_arch.py:246) is not evaluated."""

    def __init__(self, resolution=None, n_embed=8192, z_channels=3, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4),
                 num_res_blocks=2, attn_resolutions=(), embed_dim=3, beta=0.25, **ignored):
        super().__init__()
        shapes = synth.state_shapes("vqgan")
        self.encoder = ParamModule(_sub(shapes, "encoder."))
        self.decoder = ParamModule(_sub(shapes, "decoder."))
        self.quantize = VectorQuantizer2(n_embed, embed_dim, beta=beta)
        self.quant_conv = ParamModule(_sub(shapes, "quant_conv."))
        self.post_quant_conv = ParamModule(_sub(shapes, "post_quant_conv."))
        self.conv_semantic = ParamModule(_sub(shapes, "conv_semantic."))
        self._engine = None

    def _eng(self):
        key = _fingerprint(self)
        if self._engine is None or self._engine[0] != key:
            self._engine = (key, GlareEngine({}, self.state_dict(), device=next(self.parameters()).device, flow=False))
        return self._engine[1]

    @torch.no_grad()
    def encode(self, x):
        eng = self._eng()
        return eng.run_verified(lambda: eng.vqgan_encode(x)), None

    @torch.no_grad()
    def decode(self, h, vgg_feat=None):
        quant, emb_loss, _ = self.quantize(h)
        eng = self._eng()
        return None, emb_loss, eng.run_verified(lambda: eng.vq_decoder_features(quant))


class VQLLFLOWDeformable(nn.Module):
    """VQLLFLOWDeformable_arch.py:18-250.  forward(net_vq=..., lr=..., reverse=True) -> (rec_deformable, enc_feat);
    forward(gt=..., lr=..., reverse=False) -> (z, nll, logdet)  [LLFlowVQGAN2_arch.py:75-122 objective, forward only]."""

    def __init__(self, in_nc=3, out_nc=3, nf=64, nb=23, gc=32, scale=1, K=12, opt=None, step=None, which="netG",
                 fix_modules=("RRDB", "flowUpsamplerNet")):
        super().__init__()
        shapes = synth.state_shapes(which)
        self.opt = opt
        self.RRDB = ParamModule(_sub(shapes, "RRDB."))
        self.flowUpsamplerNet = FlowUpsamplerNet((80, 80, 3), 64, K, opt=opt, shapes=_sub(shapes, "flowUpsamplerNet."))
        try:
            self.quant = 255 if opt["datasets"]["train"]["quant"] is None else opt["datasets"]["train"]["quant"]   # VQLLFLOWDeformable_arch.py:28
        except (TypeError, KeyError):
            self.quant = 255
        if which == "netG":
            self.deformable_decoder = ParamModule(_sub(shapes, "deformable_decoder."))
            for name in (fix_modules or ()):                              # VQLLFLOWDeformable_arch.py:49-52: stage 3 trains the decoder only
                for prm in getattr(self, name).parameters():
                    prm.requires_grad = False
        self.dense_name = "auto"
        self._engine = None
        self._frozen = None
        self._train_ctx = None
        # True: stage-2 training calls replay as one CUDA graph per (shapes, mean branch): 106 -> 97 ms per step at batch 4 x 320x320 with the
        # final kernels (nothing while the step was 190 ms of kernels, profiles/r48_*).  Off by default: the gradients handed to autograd are
        # then the graph's static buffers (valid until the next training call) and every branch holds a second copy of the tape's memory
        self.train_graph = False
        self._train_graphs = {}
        # True: the frozen stages (encoder, flow, VQGAN) of a stage-3 training call replay as one CUDA graph per input shape.  Off by default:
        # measured on B200 at batch 2 x 256x256 the step is 62.6 ms either way (the GPU, not the launch rate, bounds those stages)
        self.stage3_graph = False

    def engine(self, net_vq=None):
        key = _fingerprint(self) + (_fingerprint(net_vq) if net_vq is not None else ())
        if self._engine is None or self._engine[0] != key:
            from .dense import make_dense
            sd_v = net_vq.state_dict() if net_vq is not None else {}
            eng = GlareEngine(self.state_dict(), sd_v, device=next(self.parameters()).device, dense=make_dense(self.dense_name),
                              decoders=net_vq is not None)
            self._engine = (key, eng)
        return self._engine[1]

    def forward(self, net_vq=None, gt=None, lr=None, z=None, eps_std=None, reverse=True, epses=None, reverse_with_grad=True,
                lr_enc=None, add_gt_noise=False, step=None, y_label=None, align_condition_feature=False, get_color_map=False):
        if get_color_map:
            raise NotImplementedError("get_color_map is not reachable from the shipped configurations")
        if reverse:
            assert lr.shape[1] == 3
            if (reverse_with_grad and self.training and torch.is_grad_enabled() and hasattr(self, "deformable_decoder")
                    and any(p.requires_grad for p in self.deformable_decoder.parameters())):
                return self._reverse_flow_train(net_vq, lr)
            with torch.no_grad():
                st = {}
                out = self.engine(net_vq).infer(lr, stages=st)
            return out, st["z_flow"]
        if add_gt_noise:
            raise NotImplementedError("add_gt_noise=True is not reachable from the shipped configurations (LLFlow_model.py:215 passes the default)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._normal_flow_train(gt, lr)
        with torch.no_grad():
            eng = self.engine(None)
            lr_d = lr.to(eng.device, torch.float32)
            enc = eng.run_verified(lambda: eng.cond_encoder(lr_d))
            zz, logdet = eng.flow_encode(gt.to(eng.device, torch.float32), enc["cond_feat"])
            nll = flowmod.gaussian_nll(zz, enc["color_map"], logdet)
        return zz, nll, logdet

    def _leaves(self):
        from . import encoder_train
        from .dense import make_dense
        if self._train_ctx is None:
            from . import flow_train
            dense = make_dense(self.dense_name)
            kernels = flow_train.CudaKernels(mode=dense.mode if dense.mode in (0, 4) else 4)
            self._train_ctx = (dense, encoder_train.CudaLeaves(dense), kernels)
        return self._train_ctx

    def _reverse_flow_train(self, net_vq, lr):
        """training call of VQLLFLOWDModel.optimize_parameters (VQLLFLOWD_model.py:207-211): ``rec, _ = netG(net_vq=..., lr=...,
        reverse=True, reverse_with_grad=True)``.  Encoder, flow and VQGAN run without gradient (VQLLFLOWDeformable_arch.py:230-248) on an
        engine keyed on the frozen modules only -- it survives the optimizer's updates of the decoder -- and the deformable decoder is one
        autograd node over its parameters (decoder_train.DeformableDecoderFn, DCN backward inside), with the whole-batch mean ratio of :567."""
        from . import decoder_train
        frozen = [self.RRDB, self.flowUpsamplerNet, net_vq]
        key = sum((_fingerprint(m) for m in frozen), ())
        if self._frozen is None or self._frozen[0] != key:
            from .dense import make_dense
            sd = {k: v for k, v in self.state_dict().items() if k.startswith(("RRDB.", "flowUpsamplerNet."))}
            self._frozen = (key, GlareEngine(sd, net_vq.state_dict(), device=next(self.parameters()).device, dense=make_dense(self.dense_name),
                                             decoders=False))
        eng = self._frozen[1]
        z, vq_feats, mid = eng.stage3_inputs(lr, graph=self.stage3_graph)
        named = [(k, p) for k, p in self.named_parameters() if k.startswith("deformable_decoder.")]
        rec = decoder_train.deformable_decoder(named, z, vq_feats, mid, self._leaves()[1], global_ratio=True)
        return rec, z

    def _normal_flow_train(self, gt, lr):
        """training call of LLFlow_model.optimize_parameters (LLFlow_model.py:215-232): ``z, nll, _ = netG(gt=..., lr=..., reverse=False)``
        followed by ``scaler.scale(nll.mean()).backward()``.  The objective and every parameter gradient come from the library's kernels
        (encoder_train.stage2_nll: one autograd node over this module's named parameters); the `mean = gt` branch is drawn per call with
        opt['train_gt_ratio'] like LLFlowVQGAN2_arch.py:109."""
        from . import encoder_train
        from .dense import make_dense
        dev = next(self.parameters()).device
        dense, leaves, flow_kernels = self._leaves()
        ratio = 0.0
        try:
            ratio = float(self.opt["train_gt_ratio"] or 0.0)
        except (TypeError, KeyError):
            pass
        named = [(k, p) for k, p in self.named_parameters() if k.startswith(("RRDB.", "flowUpsamplerNet."))]
        nll = encoder_train.stage2_nll(named, gt.to(dev, torch.float32), lr.to(dev, torch.float32), leaves,
                                       lambda x, w: dense.conv2d(x, w).float(), flow_kernels=flow_kernels, train_gt_ratio=ratio,
                                       graph=self._train_graphs if self.train_graph else False)
        return None, nll, None
