"""Public entry: host images in, enhanced host images out (what infer_dataset_lol.py:113-136 /
infer_unpaired.py:110-136 do per image, batched and with the pre/post-processing on the device).

    enh = GlareEnhancer(sd_g, sd_vq, device="cuda:0", pad="lol")
    out_u8 = enh.enhance(images_u8)          # [B,H,W,3] uint8 RGB host tensor -> same shape/dtype

pad="lol":  reflect-pad 20 px bottom + left, crop [:h, 20:]          (infer_dataset_lol.py:124,135)
pad="auto": reflect-pad (edge-repeating, cv2.BORDER_REFLECT) to the next multiple of 16, split evenly
            (infer_unpaired.py:81-88,130)
"""
import torch

from . import ops
from .engine import GlareEngine


class GlareEnhancer:
    def __init__(self, sd_g, sd_vq, device="cuda:0", pad="lol", dense=None):
        if dense is None:
            from .dense import make_dense
            dense = make_dense("auto")          # tcgen05 dense path, fp32-grade (bf16x3 split operands)
        self.engine = GlareEngine(sd_g, sd_vq, device=device, dense=dense)
        self.device = self.engine.device
        if pad not in ("lol", "auto"):
            raise ValueError("pad must be 'lol' or 'auto'")
        self.pad = pad
        self._pin_in = self._pin_out = None

    def preprocess(self, img_u8_dev):
        """uint8 [B,H,W,3] on device -> (lr [B,3,Hp,Wp] fp32, crop box); one kernel (glare_preprocess_u8)"""
        B, h, w, _ = img_u8_dev.shape
        if self.pad == "lol":
            pad, mode = (0, 20, 20, 0), 0                                        # impad(bottom=20, left=20), np.pad 'reflect'
            box = (0, h, 20, 20 + w)
        else:
            times = 16
            h1, w1 = (times - h % times) // 2, (times - w % times) // 2
            h2, w2 = (times - h % times) - h1, (times - w % times) - w1
            pad, mode = (h1, h2, w1, w2), 1                                      # cv2.BORDER_REFLECT
            box = (h1, h1 + h, w1, w1 + w)
        return ops.preprocess_u8(img_u8_dev.contiguous(), pad, mode), box

    def postprocess(self, out, box):
        """rgb(): clip to [0,1], *255, truncate to uint8 (infer_unpaired.py:40-42), NHWC; one kernel (glare_postprocess_u8)"""
        return ops.postprocess_u8(out.float(), box)

    def _device_path(self, dev_u8):
        """uint8 [B,H,W,3] on the device -> enhanced uint8 [B,H,W,3] on the device (one pass, no host synchronisation)"""
        lr, box = self.preprocess(dev_u8)
        return self.postprocess(self.engine._forward(lr)[0], box)

    @torch.no_grad()
    def enhance_device(self, dev_u8, graph=True):
        """device-resident form of `enhance`.  graph=True: pre-processing, the whole network and post-processing replay as ONE captured
        CUDA graph per input shape (engine.graphed); the returned tensor is the graph's static output, overwritten by the next call."""
        if graph:
            return self.engine.graphed("enhance-" + self.pad, self._device_path, dev_u8)
        lr, box = self.preprocess(dev_u8)
        return self.postprocess(self.engine.infer(lr), box)

    @torch.no_grad()
    def enhance(self, images_u8, out=None, graph=True):
        """images_u8: host uint8 [B,H,W,3] (pinned for async copies).  Returns host uint8 [B,H,W,3]."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise ValueError("expected uint8 [B,H,W,3]")
        with torch.cuda.device(self.device):
            dev = images_u8.to(self.device, non_blocking=True)
            res = self.enhance_device(dev, graph=graph)
            if out is None:
                out = torch.empty(res.shape, dtype=torch.uint8, pin_memory=True)
            out.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out
