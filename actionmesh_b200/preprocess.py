"""GPU image preprocessing for the DinoV2 frame encoder — replaces the host `BitImageProcessor.preprocess` call of
actionmesh/model/image_encoder.py:48-51 (on the path of row a2).

Semantics = the reference's pinned `transformers<5` (requirements.txt:10) slow image processor, i.e. Pillow:
    convert("RGB") -> PIL bicubic resize so that the SHORTEST edge is `shortest_edge` (long edge int(short*long/short))
    -> centre crop (crop_h, crop_w) -> uint8 * (1/255) (float64 product, cast to float32) -> (x - mean) / std -> CHW.
Pillow resizes uint8 images with a two-pass separable INTEGER convolution (libImaging/Resample.c): the float64 filter
weights are normalised, converted to int32 with 22 fractional bits (round half away from zero), each pass accumulates
from 1 << 21, shifts and clamps to uint8.  `resample_table` rebuilds those tables on the host (float64, same operation
order as the C code); the two CUDA kernels (csrc/preprocess.cu) run the integer passes, restricted to the rows/columns the
crop keeps, fused with rescale + normalise + layout.  The uint8 result is bit-identical to PIL's.

Host->device traffic is the raw uint8 frames (T*H*W*3 B) instead of the fp32 crops.  There is no CPU fallback: a CPU
device is an error.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._lib import AmbError

PRECISION_BITS = 32 - 8 - 2  # Resample.c: PRECISION_BITS
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bicubic(x: np.ndarray) -> np.ndarray:
    """Resample.c bicubic_filter, a = -0.5 (float64)."""
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


@lru_cache(maxsize=64)
def resample_table(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BICUBIC filter and the full-image box.
    Returns bounds (out_size, 2) int32 = (first source index, tap count) and coefficients (out_size, ksize) int32."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast == trunc; args are > -1 here
    xmin = np.where(center - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    taps = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((taps + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(taps < xmax[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for x in range(ksize):                      # sequential accumulation in tap order, like the C loop
        ww = ww + w[:, x]
    k = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    v = k * float(1 << PRECISION_BITS)
    kk = np.where(k < 0, np.trunc(-0.5 + v), np.trunc(0.5 + v)).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kk


def resize_output_size(height: int, width: int, shortest_edge: int) -> Tuple[int, int]:
    """transformers 4.x image_transforms.get_resize_output_image_size(size=int, default_to_square=False) -> (h, w)."""
    short, long = (width, height) if width <= height else (height, width)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if width <= height else (new_short, new_long)


class B200ImagePreprocessor:
    """Callable stand-in for `BitImageProcessor.preprocess(images, return_tensors="pt").pixel_values` on the GPU."""

    def __init__(self, shortest_edge: int = 256, crop_size: Tuple[int, int] = (224, 224), rescale_factor: float = 1 / 255.0,
                 image_mean: Sequence[float] = IMAGENET_MEAN, image_std: Sequence[float] = IMAGENET_STD):
        self.shortest_edge, self.crop_size = int(shortest_edge), (int(crop_size[0]), int(crop_size[1]))
        self.image_mean = tuple(float(np.float32(m)) for m in image_mean)
        self.image_std = tuple(float(np.float32(s)) for s in image_std)
        # rescale(): image.astype(float64) * scale, cast to float32 -> a 256-entry table
        self._lut_host = (np.arange(256, dtype=np.float64) * float(rescale_factor)).astype(np.float32)
        self._dev_cache: dict = {}

    @classmethod
    def from_hf(cls, proc) -> "B200ImagePreprocessor":
        """Take the numbers of an HF BitImageProcessor config (whatever backend that object itself would use)."""
        def get(obj, key):
            if isinstance(obj, dict):
                return obj.get(key)
            try:
                return obj[key]
            except (KeyError, TypeError, IndexError):
                return getattr(obj, key, None)

        short = get(proc.size, "shortest_edge")
        if short is None:
            raise AmbError(f"only shortest_edge resizing is supported (got {proc.size})")
        if int(getattr(proc, "resample", 3)) != 3:
            raise AmbError("only bicubic resampling (PIL.Image.BICUBIC == 3) is supported")
        for flag in ("do_resize", "do_center_crop", "do_rescale", "do_normalize"):
            if not getattr(proc, flag, True):
                raise AmbError(f"unsupported preprocessor config: {flag}=False")
        return cls(short, (get(proc.crop_size, "height"), get(proc.crop_size, "width")), proc.rescale_factor,
                   proc.image_mean, proc.image_std)

    def _plan(self, H: int, W: int, dev: torch.device):
        key = (H, W, str(dev))
        if key in self._dev_cache:
            return self._dev_cache[key]
        oh, ow = resize_output_size(H, W, self.shortest_edge)
        ch, cw = self.crop_size
        if oh < ch or ow < cw:
            raise AmbError(f"crop {self.crop_size} larger than the resized image {(oh, ow)}")
        top, left = (oh - ch) // 2, (ow - cw) // 2
        # Pillow skips a pass whose size does not change; an identity table (one tap of weight 1.0) is bit-equivalent
        def table(n_in, n_out):
            if n_in == n_out:
                b = np.stack([np.arange(n_out), np.ones(n_out)], axis=1).astype(np.int32)
                return b, np.full((n_out, 1), 1 << PRECISION_BITS, dtype=np.int32)
            return resample_table(n_in, n_out)

        bh, kh = table(W, ow)
        bv, kv = table(H, oh)
        bh, kh = bh[left:left + cw], kh[left:left + cw]
        bv, kv = bv[top:top + ch], kv[top:top + ch]
        y0 = int(bv[:, 0].min())
        y1 = int((bv[:, 0] + bv[:, 1]).max())
        assert 0 <= y0 < y1 <= H and int(bh[:, 0].min()) >= 0 and int((bh[:, 0] + bh[:, 1]).max()) <= W
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        plan = dict(y0=y0, n_rows=y1 - y0, bh=to(bh), kh=to(kh), bv=to(bv), kv=to(kv), lut=to(self._lut_host))
        self._dev_cache[key] = plan
        return plan

    @torch.no_grad()
    def preprocess_u8(self, frames: torch.Tensor, device, return_u8: bool = False):
        """frames (n, H, W, 3|4) uint8 (host or device) -> pixel_values (n, 3, crop_h, crop_w) fp32 on `device`."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise AmbError("B200ImagePreprocessor only runs on a CUDA (sm_100) device; there is no CPU path")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] not in (3, 4):
            raise AmbError(f"frames must be (n, H, W, 3|4) uint8, got {tuple(frames.shape)} {frames.dtype}")
        x = frames.contiguous()
        x = (x.pin_memory() if not x.is_cuda else x).to(dev, non_blocking=True)
        n, H, W, _ = x.shape
        p = self._plan(H, W, dev)
        ch, cw = self.crop_size
        mid = torch.empty(n, p["n_rows"], cw, 3, dtype=torch.uint8, device=dev)
        ops.resize_h_u8(x, p["y0"], p["n_rows"], p["bh"], p["kh"], mid)
        out = torch.empty(n, 3, ch, cw, dtype=torch.float32, device=dev)
        u8 = torch.empty(n, ch, cw, 3, dtype=torch.uint8, device=dev) if return_u8 else None
        ops.resize_v_normalize(mid, p["y0"], p["bv"], p["kv"], p["lut"], self.image_mean, self.image_std, out, u8)
        return (out, u8) if return_u8 else out

    def preprocess(self, images: List, device) -> torch.Tensor:
        """images: list of PIL images (any mode; converted with .convert("RGB") like do_convert_rgb) -> (T,3,h,w) fp32.
        Frames of different sizes are processed per size group; the output keeps the input order."""
        arrs = [np.asarray(im.convert("RGB") if getattr(im, "mode", "RGB") != "RGB" else im, dtype=np.uint8) for im in images]
        out = [None] * len(arrs)
        groups: dict = {}
        for i, a in enumerate(arrs):
            groups.setdefault(a.shape, []).append(i)
        for idx in groups.values():
            pv = self.preprocess_u8(torch.from_numpy(np.stack([arrs[i] for i in idx])), device)
            for j, i in enumerate(idx):
                out[i] = pv[j]
        return torch.stack(out) if len(groups) > 1 else pv


class B200FramePreprocessor:
    """`ImagePreprocessor` (actionmesh/preprocessing/image_processor.py:104-146) on the GPU: RGBA frames are composited on a
    white background, cropped to the foreground bounding box (shared across the clip unless `independent_cropping`) and padded
    to a square with a `padding_ratio` margin.  Same constructor fields, same `process_images(frames) -> list[PIL.Image]`, same
    ValueError for frames without a usable alpha channel; `process_to_u8` keeps the result on the device for the encoder.

    Two kernels (amb_alpha_stats, amb_composite_crop_pad) around the host integers the reference computes too (bounding boxes,
    paddings, image_processor.py:57-101); the uint8 output is bit-identical to the reference's."""

    def __init__(self, independent_cropping: bool = False, padding_ratio: float = 0.1, device="cuda"):
        self.independent_cropping = independent_cropping
        self.padding_ratio = padding_ratio
        self.bg_color = np.array([1.0, 1.0, 1.0])
        self.device = torch.device(device)

    @staticmethod
    def _padding(w: int, h: int, padding_ratio: float):
        max_dim = max(w, h)                                 # apply_padding, image_processor.py:91-96
        pad_base = int(max_dim * padding_ratio)
        return pad_base + (max_dim - w) // 2, pad_base + (max_dim - h) // 2

    def process_to_u8(self, frames: List) -> List[torch.Tensor]:
        """-> one (H', W', 3) uint8 CUDA tensor per frame (all the same size unless independent_cropping)."""
        if self.device.type != "cuda":
            raise AmbError("B200FramePreprocessor only runs on a CUDA (sm_100) device; there is no CPU path")
        arrs = [np.ascontiguousarray(f if getattr(f, "mode", "RGBA") == "RGBA" else f.convert("RGBA")) for f in frames]
        if len({a.shape for a in arrs}) != 1:
            raise AmbError("B200FramePreprocessor: all frames of a clip must have the same size")
        with torch.cuda.device(self.device):
            rgba = torch.from_numpy(np.stack(arrs)).pin_memory().to(self.device, non_blocking=True)
            n, H, W, _ = rgba.shape
            stats = ops.alpha_stats(rgba).cpu().tolist()     # 5 integers per frame: the one host round trip of the stage
            min_count = int(H * W * 0.01)                     # is_valid_alpha(min_ratio=0.01, threshold=127), :15-23
            boxes = []
            for xmin, ymin, xmax, ymax, fg in stats:
                if not (H * W - fg >= min_count and fg >= min_count):
                    raise ValueError("Invalid alpha channel: insufficient foreground/background")
                boxes.append((xmin, ymin, xmax - xmin + 1, ymax - ymin + 1))
            if not self.independent_cropping:                # aggregate_bboxes, :69-77
                x0, y0 = min(b[0] for b in boxes), min(b[1] for b in boxes)
                x1, y1 = max(b[0] + b[2] for b in boxes), max(b[1] + b[3] for b in boxes)
                box = (x0, y0, x1 - x0, y1 - y0)
                px, py = self._padding(box[2], box[3], self.padding_ratio)
                return list(ops.composite_crop_pad(rgba, box, px, py))
            out = []
            for i, box in enumerate(boxes):
                px, py = self._padding(box[2], box[3], self.padding_ratio)
                out.append(ops.composite_crop_pad(rgba[i:i + 1], box, px, py)[0])
            return out

    def process_images(self, frames: List) -> List:
        from PIL import Image

        return [Image.fromarray(t.cpu().numpy()) for t in self.process_to_u8(frames)]
