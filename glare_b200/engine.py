"""GLARE inference engine on one B200: the hot path of VQLLFLOWDeformable.reverse_flow
(VQLLFLOWDeformable_arch.py:222-250), state-dict driven, kernels from libglare_b200.so.

    cond-encoder (ConditionEncoder.py:46-55) -> inverse flow (FlowUpsamplerNet.py:290-326) -> VQ lookup
    (quantize.py:271-312) -> VQGAN decoder features (VQModel_arch.py:81-91) -> AFT decoder with DCNv2 warps
    (deformableDecoder_arch.py:525-576)

The engine consumes the reference's checkpoints unchanged (``net_G.pth`` / ``vqgan.pkl`` state-dict keys).
It is CUDA-only by construction: there is no CPU path and no PyTorch re-implementation of the VQ, flow-step
or DCN operators to fall back to.

Dense operators (3x3 / 1x1 convolutions, GroupNorm+swish, attention) go through the ``dense`` backend object
(``glare_b200.dense.TcDense``: tcgen05 implicit-GEMM kernels of libglare_b200.so).  There is no library (cuDNN / cuBLAS)
backend in the package; the cuDNN fp32 comparison path of the tests lives in ``tests/libdense.py``.
"""
import torch
import torch.nn.functional as F

from . import flow as flowmod
from . import ops


class _Timed:
    """CUDA-event bracket on the current stream around one named operator (used by bench.py for the roofline)."""

    def __init__(self, timers, name):
        self.timers, self.name = timers, name

    def __enter__(self):
        if self.timers is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.timers is not None:
            self.e1.record()
            self.timers.setdefault(self.name, []).append((self.e0, self.e1))


class GlareEngine:
    def __init__(self, sd_g, sd_vq, device="cuda:0", dense=None, per_sample_ratio=True, flow=True, decoders=True):
        if not torch.cuda.is_available():
            raise RuntimeError("glare_b200.GlareEngine needs a CUDA device (B200 / sm_100a); there is no CPU path")
        self.device = torch.device(device)
        if dense is None:
            from .dense import make_dense
            dense = make_dense("auto")                     # tcgen05, fp32-grade operands
        self.dense = dense
        self.per_sample_ratio = per_sample_ratio
        self.timers = None        # bench.py sets a dict: name -> [(start_event, end_event), ...]
        self._graphs = {}
        self.g = {k: v.to(self.device, torch.float32).contiguous() for k, v in sd_g.items()
                  if not k.startswith(("flowUpsamplerNet.f.", "deformable_decoder.scale", "deformable_decoder.bias",
                                       "deformable_decoder.enc", "deformable_decoder.conv_out"))}
        self.v = {k: v.to(self.device, torch.float32).contiguous() for k, v in sd_vq.items()
                  if k.startswith(("quantize.", "post_quant_conv.", "decoder.", "encoder.", "quant_conv."))}
        with torch.cuda.device(self.device):
            self.flow_plan = flowmod.FlowPlan(sd_g, self.device) if flow and "flowUpsamplerNet.layers.0.actnorm.bias" in sd_g else None
            self.codebook_packed = (ops.vq_pack_codebook(self.v["quantize.embedding.weight"])
                                    if "quantize.embedding.weight" in self.v else None)
            self._one = torch.ones((1,), device=self.device, dtype=torch.float32)
            self.dcn_w = {i: ops.dcn_pack_weight(self.g["deformable_decoder.warp.%d.dcn.weight" % i]) for i in (0, 1)
                          if decoders and ("deformable_decoder.warp.%d.dcn.weight" % i) in self.g}

    def _timed(self, name):
        return _Timed(self.timers, name)

    # ------------------------------------------------------------------ taming blocks (encoder_decoder.py)
    def _conv(self, sd, p, x, stride=1, padding=1, residual=None):
        return self.dense.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding, residual=residual)

    def _gn(self, sd, p, x, swish=True):
        return self.dense.gn_swish(x, sd[p + ".weight"], sd[p + ".bias"], swish)

    def resnet_block(self, sd, p, x):
        """ResnetBlock.forward   encoder_decoder.py:117-137"""
        h = self._conv(sd, p + ".conv1", self._gn(sd, p + ".norm1", x))
        if (p + ".nin_shortcut.weight") in sd:
            x = self._conv(sd, p + ".nin_shortcut", x, padding=0)
        return self._conv(sd, p + ".conv2", self._gn(sd, p + ".norm2", h), residual=x)      # x + h fused in the epilogue

    def attn_block(self, sd, p, x):
        """AttnBlock.forward   encoder_decoder.py:168-192"""
        hn = self._gn(sd, p + ".norm", x, swish=False)
        q = k = None
        if getattr(self.dense, "attn_fused", False) and hasattr(self.dense, "conv2d_operand"):
            # q / k only feed the scores GEMM: their convs write its operands directly (and, for the Cauchy-Schwarz softmax reference, the
            # row norms)
            rsq = getattr(self.dense, "attn_ref", "sampled") == "cauchy"
            q = self.dense.conv2d_operand(hn, sd[p + ".q.weight"], sd.get(p + ".q.bias"), row_sq=rsq)
            k = self.dense.conv2d_operand(hn, sd[p + ".k.weight"], sd.get(p + ".k.bias"), row_sq=rsq)
        if q is None or k is None:
            q = self._conv(sd, p + ".q", hn, padding=0)
            k = self._conv(sd, p + ".k", hn, padding=0)
        v = self._conv(sd, p + ".v", hn, padding=0)
        if hasattr(self.dense, "conv2d_operand"):
            o = self.dense.attention(q, k, v, as_operand=True)           # proj_out's operand straight from the P V epilogue when fused
        else:
            o = self.dense.attention(q, k, v)
        return self._conv(sd, p + ".proj_out", o, padding=0, residual=x)

    def downsample(self, sd, p, x):
        """encoder_decoder.py:68-72"""
        if hasattr(self.dense, "downsample_conv"):
            y = self.dense.downsample_conv(x, sd[p + ".conv.weight"], sd.get(p + ".conv.bias"))
            if y is not None:
                return y
        return self._conv(sd, p + ".conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)

    def upsample(self, sd, p, x):
        """encoder_decoder.py:49-53"""
        if hasattr(self.dense, "upsample_conv"):
            y = self.dense.upsample_conv(x, sd[p + ".conv.weight"], sd.get(p + ".conv.bias"))
            if y is not None:
                return y
        return self._conv(sd, p + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))

    def vqgan_encode(self, x):
        """VQModel.encode (VQModel_arch.py:74-79): Encoder + quant_conv"""
        x = x.to(self.device, torch.float32)
        return self._conv(self.v, "quant_conv", self._encoder(self.v, "encoder", x)[0], padding=0).float()

    def cond_encoder(self, x, p="RRDB"):
        """ConEncoder1.forward   ConditionEncoder.py:46-55"""
        sd = self.g
        enc, mid = self._encoder(sd, p + ".encoder", x)
        return {"cond_feat": torch.sigmoid(self._conv(sd, p + ".cond_conv.0", enc).float()),
                "color_map": self._conv(sd, p + ".color_conv", enc).float(), "mid_feat": mid}

    def _encoder(self, sd, e, x):
        """Encoder.forward(mid_feat=True)   encoder_decoder.py:406-442"""
        h = self._conv(sd, e + ".conv_in", x)
        mid = []
        for lvl in range(3):
            for blk in range(2):
                h = self.resnet_block(sd, "%s.down.%d.block.%d" % (e, lvl, blk), h)
                if ("%s.down.%d.attn.%d.q.weight" % (e, lvl, blk)) in sd:
                    h = self.attn_block(sd, "%s.down.%d.attn.%d" % (e, lvl, blk), h)
            if lvl != 2:
                mid.append(h)
                h = self.downsample(sd, "%s.down.%d.downsample" % (e, lvl), h)
        h = self.resnet_block(sd, e + ".mid.block_1", h)
        h = self.attn_block(sd, e + ".mid.attn_1", h)
        h = self.resnet_block(sd, e + ".mid.block_2", h)
        return self._conv(sd, e + ".conv_out", self._gn(sd, e + ".norm_out", h)).float(), mid

    def _decoder_trunk(self, sd, p, z):
        h = self._conv(sd, p + ".conv_in", z)
        h = self.resnet_block(sd, p + ".mid.block_1", h)
        h = self.attn_block(sd, p + ".mid.attn_1", h)
        return self.resnet_block(sd, p + ".mid.block_2", h)

    def vq_decoder_features(self, zq, p="decoder"):
        """VQModel.decode after the quantizer (VQModel_arch.py:88-90, encoder_decoder.py:515-551)"""
        sd = self.v
        h = self._decoder_trunk(sd, p, self._conv(sd, "post_quant_conv", zq, padding=0))
        feats = []
        for lvl in (2, 1, 0):
            for blk in range(3):
                h = self.resnet_block(sd, "%s.up.%d.block.%d" % (p, lvl, blk), h)
                if lvl == 2:
                    h = self.attn_block(sd, "%s.up.%d.attn.%d" % (p, lvl, blk), h)
            if lvl != 2:
                feats.append(h)
            if lvl != 0:
                h = self.upsample(sd, "%s.up.%d.upsample" % (p, lvl), h)
        return feats

    @staticmethod
    def _can_fuse_axpby(a, b):
        return (a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape == b.shape and a.stride() == b.stride() and a[0].numel() % 4 == 0
                and (a.is_contiguous() or a.is_contiguous(memory_format=torch.channels_last)))

    # ------------------------------------------------------------------ AFT / DCN (deformableDecoder_arch.py)
    def warp_block(self, i, x_vq, h):
        """WarpBlock.forward (:285-290) + DCNv2Pack.forward (:141-152)"""
        sd, p = self.g, "deformable_decoder.warp.%d" % i
        cat = self.dense.cat_operand(x_vq, h) if hasattr(self.dense, "cat_operand") else None     # the conv operand of cat([x_vq, h]) in one pass
        if cat is None:
            cat = torch.cat([x_vq, h], dim=1)
        feat = None
        if hasattr(self.dense, "conv2d_operand"):                        # feat only feeds conv_offset: written as its operand
            feat = self.dense.conv2d_operand(cat, sd[p + ".offset.weight"], sd.get(p + ".offset.bias"))
        if feat is None:
            feat = self._conv(sd, p + ".offset", cat)
        out = self._conv(sd, p + ".dcn.conv_offset", feat).float()
        if hasattr(self.dense, "dcn_pack"):
            y = self.dense.dcn_pack(x_vq.float(), out, sd[p + ".dcn.weight"], sd[p + ".dcn.bias"], 4)
            if y is not None:
                return y
        o1, o2, m = torch.chunk(out, 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        mask = torch.sigmoid(m)
        x_vq = x_vq.float().contiguous()
        with self._timed("dcn%d" % i):
            return ops.modulated_deform_conv(x_vq, offset, mask, sd[p + ".dcn.weight"], sd[p + ".dcn.bias"], 1, 1, 1, 1, 4,
                                             packed_weight=self.dcn_w[i])

    def aft_decoder(self, z, vq_feats, enc_feats, p="deformable_decoder"):
        """MultiScaleDecoder2.forward   deformableDecoder_arch.py:525-576"""
        sd = self.g
        h = self._decoder_trunk(sd, p, z)
        for lvl in (2, 1, 0):
            for blk in range(3):
                h = self.resnet_block(sd, "%s.up.%d.block.%d" % (p, lvl, blk), h)
                if lvl == 2:
                    h = self.attn_block(sd, "%s.up.%d.attn.%d" % (p, lvl, blk), h)
            if lvl != 2:
                mixf = torch.sigmoid(sd["%s.mix.%d.w" % (p, 1 - lvl)])
                fused = self._can_fuse_axpby(enc_feats[lvl], h) and mixf.numel() == 1
                if fused:
                    h = ops.aft_axpby(enc_feats[lvl], h, mixf, 1 - mixf)         # Mix.forward :587-590 in one pass
                else:
                    h = enc_feats[lvl] * mixf + h * (1 - mixf)
                x_vq = self.warp_block(1 - lvl, vq_feats[1 - lvl], h).to(h.dtype)
                if self.per_sample_ratio:
                    # :567 reduces over the whole batch but the reference only ever runs batch 1; per-sample
                    # means keep that behaviour for any batch size / sharding (DESIGN.md "batch coupling")
                    ratio = h.float().mean(dim=(1, 2, 3), keepdim=True) / x_vq.float().mean(dim=(1, 2, 3), keepdim=True)
                else:
                    ratio = h.float().mean() / x_vq.float().mean()
                if self.per_sample_ratio and self._can_fuse_axpby(h, x_vq):
                    h = ops.aft_axpby(h, x_vq, self._one, ratio)                 # h * 1 is exact: h + x_vq * ratio, one pass
                else:
                    h = h + x_vq * ratio.to(h.dtype)
            if lvl != 0:
                h = self.upsample(sd, "%s.up.%d.upsample" % (p, lvl), h)
        return self._conv(sd, p + ".residual_conv", self._gn(sd, p + ".norm_out", h)).float()

    # ------------------------------------------------------------------ stages
    def flow_decode(self, z, ft, trace=None):
        with self._timed("flow_total"):
            return flowmod.decode(self.flow_plan, z, ft, lambda x, w: self.dense.conv2d(x, w).float(), trace=trace)[0]

    def flow_encode(self, gt, ft, logdet=None):
        return flowmod.encode(self.flow_plan, gt, ft, lambda x, w: self.dense.conv2d(x, w).float(), logdet=logdet)

    def vector_quantize(self, z):
        with self._timed("vq"):
            idx, zq = ops.vq_lookup(z, self.codebook_packed)
        return zq, idx

    def _forward(self, lr):
        """one pass of the hot path on device tensors; returns (out, stage tensors)"""
        enc = self.cond_encoder(lr)
        z = self.flow_decode(enc["color_map"], enc["cond_feat"])
        zq, idx = self.vector_quantize(z)
        vq_feats = self.vq_decoder_features(zq)
        out = self.aft_decoder(z, vq_feats, enc["mid_feat"])
        return out, dict(cond_feat=enc["cond_feat"], color_map=enc["color_map"], mid0=enc["mid_feat"][0], mid1=enc["mid_feat"][1],
                         z_flow=z, z_q=zq, idx=idx, vq_feat1=vq_feats[0], vq_feat0=vq_feats[1], out=out)

    @torch.no_grad()
    def stage3_inputs(self, lr, graph=False):
        """the frozen part of VQLLFLOWDeformable.reverse_flow (VQLLFLOWDeformable_arch.py:230-248, all under no_grad there): condition encoder
        -> flow decode -> VQ -> VQGAN decoder features.  -> (z [B,3,h,w], [vq feat @2h, vq feat @4h], encoder mid features by level).
        graph=True: one CUDA-graph replay per input shape (~600 launches at a training crop's size, CPU-launch bound otherwise); the results
        are copied out of the graph's static buffers, so they stay valid across later calls (a training tape holds them until its backward)."""
        with torch.cuda.device(self.device):
            lr = lr.to(self.device, torch.float32)

            def run(x):
                enc = self.cond_encoder(x)
                z = self.flow_decode(enc["color_map"], enc["cond_feat"])
                zq, _ = self.vector_quantize(z)
                return z, self.vq_decoder_features(zq), [enc["mid_feat"][0], enc["mid_feat"][1]]

            if graph:
                z, vq, mid = self.graphed("stage3_inputs", run, lr)
                return z.clone(), [t.clone() for t in vq], [t.clone() for t in mid]
            return self.run_verified(lambda: run(lr))

    def _verified(self):
        # the fused-softmax attention (dense.py) verifies its row sums on the device; a tripped flag switches the backend to the exact
        # softmax path and the caller recomputes (4-byte read, once per call)
        verified = getattr(self.dense, "attention_verified", None)
        return verified is None or verified()

    def run_verified(self, fn):
        """run ``fn()`` (any part of the path that contains AttnBlocks) with the fused-softmax safety net of `infer`: recompute once on the
        exact softmax path if the device flag tripped"""
        out = fn()
        return out if self._verified() else fn()

    @torch.no_grad()
    def infer(self, lr, stages=None, graph=False):
        """lr [B,3,H,W] = log(clamp(x + 1e-3)) (infer_unpaired.py:121-122), H and W multiples of 4; returns RGB [B,3,H,W] fp32.
        graph=True: the whole pass is ONE captured CUDA graph per input shape (see `graphed`); the returned tensor (and `stages`) are the
        graph's static buffers, overwritten by the next call with the same shape."""
        if lr.dim() != 4 or lr.shape[1] != 3 or lr.shape[2] % 4 or lr.shape[3] % 4:
            # the encoder halves the resolution twice and the decoders double it back onto the encoder's skip features
            # (deformableDecoder_arch.py:553-566); the reference entry points pad to multiples of 16 / by 20 for this reason
            raise ValueError("expected lr [B,3,H,W] with H and W multiples of 4 (pad first, see api.GlareEnhancer); got %s" % (tuple(lr.shape),))
        with torch.cuda.device(self.device):
            lr = lr.to(self.device, torch.float32)
            if graph:
                out, st = self.graphed("infer", self._forward, lr)
            else:
                for _ in range(2):
                    out, st = self._forward(lr)
                    if self._verified():
                        break
        if stages is not None:
            stages.update(st)
        return out

    # ------------------------------------------------------------------ whole-pass CUDA graphs
    def graphed(self, name, fn, x):
        """Run ``fn(x)`` (device tensor in, tensors out) as one CUDA graph captured per (name, shape, dtype): one host launch per pass instead of
        ~870 ctypes launches.  Every kernel of libglare_b200.so takes its stream as an argument and keeps no host state, TMA descriptors are
        passed by value in the kernel parameters, outputs come from torch's caching allocator -- so the pass captures as is.  The attention
        safety flag is read after every replay; if it trips, the backend leaves the fused-softmax path, every graph is dropped and the
        pass is re-captured and re-run (same semantics as the eager path)."""
        key = (name, tuple(x.shape), x.dtype)
        for _ in range(2):
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = _Graphed(self, fn, x)
            out = g(x)
            if self._verified():
                return out
            self._graphs.clear()
        raise RuntimeError("attention safety flag tripped on the exact softmax path")


class _Graphed:
    """``fn(*tensors)`` captured as one CUDA graph: static input copies, an eager warm-up pass on a side stream (weight-packing caches,
    lazily created buffers, backend mode), then the capture.  ``verify()`` (optional) is polled during the warm-up: False means "run it
    again" (the attention backend just switched paths).  ``capture_ctx`` (optional context manager) wraps the captured pass only."""

    def __init__(self, engine, fn, *examples, verify=None, capture_ctx=None):
        import contextlib
        from . import ops
        self.ops = ops
        self.static_in = [e.clone() for e in examples]
        if verify is None and engine is not None:
            verify = engine._verified
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                fn(*self.static_in)
                if verify is None or verify():
                    break
        cur.wait_stream(side)
        n0 = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with (capture_ctx() if capture_ctx is not None else contextlib.nullcontext()):
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.static_out = fn(*self.static_in)
        self.kernels = ops.LAUNCHES - n0               # kernels of libglare_b200.so inside the graph (ops.LAUNCHES is advanced per replay)

    def __call__(self, *xs):
        for dst, x in zip(self.static_in, xs):
            dst.copy_(x, non_blocking=True)
        self.graph.replay()
        self.ops.LAUNCHES += self.kernels
        return self.static_out
