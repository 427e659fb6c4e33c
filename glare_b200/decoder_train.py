"""Stage-3 training of the deformable decoder on the library's kernels: MultiScaleDecoder2.forward (deformableDecoder_arch.py:525-576) with a
tape and the backward pass of every block -- what the reference gets from torch autograd in VQLLFLOWDModel.optimize_parameters
(VQLLFLOWD_model.py:187-232) around VQLLFLOWDeformable.reverse_flow (VQLLFLOWDeformable_arch.py:222-250): the condition encoder, the flow and
the VQGAN run under no_grad there (fix_modules, :19,49-52), only `deformable_decoder.*` receives gradients, and this is where the DCN
backward kernels (csrc/dcn_bwd.cu) run inside the loop.

Layering as in encoder_train.py: kernel-level ``Leaves`` (tensor-core conv / GEMM path, GroupNorm, softmax, DCN forward + backward kernels),
block-level backward rules in plain Python shared by every backend (tests/test_decoder_train_cpu.py runs this logic on torch leaves against
torch autograd of the oracle and of the reference's own module).
"""
import torch

from .encoder_train import BlockGraph


class DecoderTrainer(BlockGraph):
    """forward with tape and backward of MultiScaleDecoder2; parameter names are the reference's state-dict keys under ``prefix``.
    ``global_ratio``: the reference's `h.mean() / x_vq.mean()` (:567) reduces over the WHOLE batch tensor -- that is what training sees at
    batch_size 2 (train_stage3_LOL.yml:39); False gives the per-sample means the inference engine uses for batch-size independence."""

    def __init__(self, leaves, sd, prefix="deformable_decoder", global_ratio=True, dg=4):
        super().__init__(leaves, sd)
        self.p, self.global_ratio, self.dg = prefix, global_ratio, dg

    # ------------------------------------------------------------------ node types beyond the encoder's
    def _mix(self, key, fea, i):
        """Mix.forward   deformableDecoder_arch.py:587-590"""
        w = self.sd[key]
        m = torch.sigmoid(w.float())
        h = self.vals[i]

        def bwd(gy):
            self.tape._acc(key, ((gy * (fea - h)).sum() * m * (1 - m)).reshape(w.shape))
            return (gy * (1 - m),)

        return self._fn(fea * m + h * (1 - m), (i,), bwd)

    def _warp(self, p, x_vq, i):
        """WarpBlock.forward (:285-290) + DCNv2Pack.forward (:141-152); x_vq comes from the frozen VQGAN decoder: no gradient flows into it"""
        cx = x_vq.shape[1]
        cat = self._fn(torch.cat([x_vq, self.vals[i]], dim=1), (i,), lambda gy: (gy[:, cx:],))
        feat = self._conv(p + ".offset", cat)
        om = self._conv(p + ".dcn.conv_offset", feat)
        w, b = self.sd[p + ".dcn.weight"], self.sd[p + ".dcn.bias"]
        out = self.vals[om]
        self.offset_ids.append(om)                      # (tests teacher-force these values to separate arithmetic error from cell flips)
        n_off = out.shape[1] // 3 * 2                   # chunk(out, 3): offset = cat(o1, o2) = the first two thirds, mask = sigmoid(last third)

        def bwd(gy):
            offset, mask = out[:, :n_off].contiguous(), torch.sigmoid(out[:, n_off:]).contiguous()
            _, g_off, g_mask, g_w, g_b = self.L.dcn_bwd(x_vq, offset, mask, w, gy, self.dg)
            self.tape._acc(p + ".dcn.weight", g_w)
            self.tape._acc(p + ".dcn.bias", g_b)
            return (torch.cat([g_off, g_mask * mask * (1 - mask)], dim=1),)

        return self._fn(self.L.dcn_fwd(x_vq, out, w, b, self.dg), (om,), bwd)

    def _ratio_add(self, i, j):
        """h + x_vq * (h.mean() / x_vq.mean())   :567"""
        h, x = self.vals[i], self.vals[j]
        dims = (0, 1, 2, 3) if self.global_ratio else (1, 2, 3)
        n = float(h.numel() if self.global_ratio else h[0].numel())
        mh, mx = h.mean(dim=dims, keepdim=True), x.mean(dim=dims, keepdim=True)
        r = mh / mx

        def bwd(gy):
            s = (gy * x).sum(dim=dims, keepdim=True)
            return gy + s / (n * mx), gy * r - s * mh / (n * mx * mx)

        return self._fn(h + x * r, (i, j), bwd)

    def _upsample(self, p, i):
        """Upsample.forward   encoder_decoder.py:46-50: nearest x2, then the 3x3 conv"""
        up = self._fn(self.L.up2(self.vals[i]), (i,), lambda gy: (self.L.up2_adjoint(gy),))
        return self._conv(p + ".conv", up)

    # ------------------------------------------------------------------ forward / backward
    def forward(self, z, vq_feats, enc_feats):
        """z [B,3,h,w] (the flow's latent), vq_feats = [feat 256ch @2h, feat 128ch @4h] of the frozen VQGAN decoder, enc_feats[lvl] the
        condition encoder's mid features at level 1 (256ch @2h) and level 0 (128ch @4h).  -> reconstruction [B,3,4h,4w]"""
        self._begin(z)
        self.offset_ids = []
        p = self.p
        h = self._conv(p + ".conv_in", 0, need_gx=False)
        h = self._resnet(p + ".mid.block_1", h)
        h = self._attn(p + ".mid.attn_1", h)
        h = self._resnet(p + ".mid.block_2", h)
        for lvl in (2, 1, 0):
            for blk in range(3):
                h = self._resnet("%s.up.%d.block.%d" % (p, lvl, blk), h)
                if lvl == 2:
                    h = self._attn("%s.up.%d.attn.%d" % (p, lvl, blk), h)
            if lvl != 2:
                h = self._mix("%s.mix.%d.w" % (p, 1 - lvl), enc_feats[lvl].float(), h)
                x_vq = self._warp("%s.warp.%d" % (p, 1 - lvl), vq_feats[1 - lvl].float(), h)
                h = self._ratio_add(h, x_vq)
            if lvl != 0:
                h = self._upsample("%s.up.%d.upsample" % (p, lvl), h)
        self.out_id = self._conv(p + ".residual_conv", self._gn(p + ".norm_out", h, True))
        return self.vals[self.out_id]

    def backward(self, g_rec):
        """gradient of the reconstruction -> {state-dict key: gradient} of every parameter the forward uses (conv_out, scale.*, bias.*, enc.*
        are constructed but unused by the reference's forward: absent, like its ``param.grad is None``)"""
        self._backprop({self.out_id: g_rec.float()})
        grads = self.tape.grads
        self.release()
        return grads


class DeformableDecoderFn(torch.autograd.Function):
    """MultiScaleDecoder2.forward as ONE autograd node over its parameters: ``forward`` runs the decoder with the library's kernels and
    keeps the tape, ``backward`` turns the reconstruction's gradient into parameter gradients.  Drop-in for the
    ``rec_deformable = self.deformable_decoder(enc_feat, code_decoder_output, c_feat)`` call of VQLLFLOWDeformable_arch.py:249 when the module
    trains; its inputs come from no_grad regions in the reference (:237-248) and get no gradient here either."""

    @staticmethod
    def forward(ctx, cfg, z, vq1, vq0, mid1, mid0, *params):
        sd = {k: p.detach() for k, p in zip(cfg["keys"], params)}
        with torch.no_grad():
            tr = DecoderTrainer(cfg["leaves"], sd, prefix=cfg.get("prefix", "deformable_decoder"), global_ratio=cfg.get("global_ratio", True))
            rec = tr.forward(z, [vq1, vq0], {1: mid1, 0: mid0})
        ctx.trainer, ctx.keys = tr, cfg["keys"]
        return rec.clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rec):
        if ctx.trainer is None:
            raise RuntimeError("Trying to backward through the deformable decoder a second time: its tape is freed by the first backward pass "
                               "(the reference's optimize_parameters calls backward once per forward)")
        with torch.no_grad():
            grads = ctx.trainer.backward(g_rec)
        ctx.trainer = None
        return (None,) * 6 + tuple(grads.get(k) for k in ctx.keys)


def deformable_decoder(named_parameters, z, vq_feats, enc_feats, leaves, global_ratio=True, prefix="deformable_decoder"):
    """``named_parameters``: (state-dict key, Parameter) pairs of `deformable_decoder.*`; returns the reconstruction with the graph edge to
    every parameter, so that the reference's ``total_loss.backward()`` fills their ``.grad``"""
    named = list(named_parameters)
    cfg = {"keys": [k for k, _ in named], "leaves": leaves, "global_ratio": global_ratio, "prefix": prefix}
    return DeformableDecoderFn.apply(cfg, z, vq_feats[0], vq_feats[1], enc_feats[1], enc_feats[0], *[p for _, p in named])
