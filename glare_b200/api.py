"""Public entry: host images in, enhanced host images out (what infer_dataset_lol.py:113-136 /
infer_unpaired.py:110-136 do per image, batched and with the pre/post-processing on the device).

    enh = GlareEnhancer(sd_g, sd_vq, device="cuda:0", pad="lol")
    out_u8 = enh.enhance(images_u8)          # [B,H,W,3] uint8 RGB host tensor -> same shape/dtype

pad="lol":  reflect-pad 20 px bottom + left, crop [:h, 20:]          (infer_dataset_lol.py:124,135)
pad="auto": reflect-pad (edge-repeating, cv2.BORDER_REFLECT) to the next multiple of 16, split evenly
            (infer_unpaired.py:81-88,130)
"""
import torch
import torch.nn.functional as F

from . import synth
from .engine import GlareEngine


class GlareEnhancer:
    def __init__(self, sd_g, sd_vq, device="cuda:0", pad="lol", dense=None):
        if dense is None:
            from .dense import make_dense
            dense = make_dense("auto")          # tcgen05 dense path, fp32-grade (tf32 + 2 x bf16 cross terms)
        self.engine = GlareEngine(sd_g, sd_vq, device=device, dense=dense)
        self.device = self.engine.device
        if pad not in ("lol", "auto"):
            raise ValueError("pad must be 'lol' or 'auto'")
        self.pad = pad
        self._pin_in = self._pin_out = None

    def preprocess(self, img_u8_dev):
        """uint8 [B,H,W,3] on device -> (lr [B,3,Hp,Wp] fp32, crop box)"""
        x = img_u8_dev.permute(0, 3, 1, 2).float() / 255.0                       # t(): infer_unpaired.py:37
        B, _, h, w = x.shape
        if self.pad == "lol":
            x = F.pad(x, (20, 0, 0, 20), mode="reflect")                         # impad(bottom=20, left=20)
            box = (0, h, 20, 20 + w)
        else:
            times = 16
            h1, w1 = (times - h % times) // 2, (times - w % times) // 2
            h2, w2 = (times - h % times) - h1, (times - w % times) - w1
            iy = synth._symmetric_index(h, h1, h2).to(x.device)
            ix = synth._symmetric_index(w, w1, w2).to(x.device)
            x = x.index_select(-2, iy).index_select(-1, ix)
            box = (h1, h1 + h, w1, w1 + w)
        return torch.log(torch.clamp(x + 1e-3, min=1e-3)), box                   # infer_dataset_lol.py:127-128

    def postprocess(self, out, box):
        """rgb(): clip to [0,1], *255, truncate to uint8 (infer_unpaired.py:40-42), NHWC"""
        y0, y1, x0, x1 = box
        o = out[:, :, y0:y1, x0:x1].clamp(0, 1) * 255.0
        return o.to(torch.uint8).permute(0, 2, 3, 1).contiguous()

    @torch.no_grad()
    def enhance(self, images_u8, out=None):
        """images_u8: host uint8 [B,H,W,3] (pinned for async copies).  Returns host uint8 [B,H,W,3]."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise ValueError("expected uint8 [B,H,W,3]")
        with torch.cuda.device(self.device):
            dev = images_u8.to(self.device, non_blocking=True)
            lr, box = self.preprocess(dev)
            res = self.postprocess(self.engine.infer(lr), box)
            if out is None:
                out = torch.empty(res.shape, dtype=torch.uint8, pin_memory=True)
            out.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out
