"""TEST INFRASTRUCTURE ONLY — CPU restatement of the image preprocessing in front of DinoV2
(actionmesh/model/image_encoder.py:48-51 -> HF `BitImageProcessor.preprocess`).

The arithmetic lives in third-party dependencies that are NOT under /root/reference:
  * transformers, pinned `transformers<5` (requirements.txt:10; this container has 5.5.0 whose BitImageProcessor is a
    torchvision backend with different resampling): the 4.x slow processor is restated here from its published code
    (models/bit/image_processing_bit.py::preprocess, image_transforms.py::resize/center_crop/rescale/normalize);
  * Pillow (`Image.resize(..., resample=BICUBIC)`), which IS installed: `bit_preprocess_pil` calls it directly, and
    `pil_bicubic_resize_u8` restates libImaging/Resample.c in numpy; tests/test_preprocess_cpu.py pins the restatement
    bit-exactly against Pillow itself.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc (scalar loops, float64 like the C doubles)."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds, kk = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            k[x] = _bicubic((x + xmin - center + 0.5) * ss)
            ww += k[x]
        if ww != 0.0:
            for x in range(xmax):
                k[x] /= ww
        bounds.append((xmin, xmax))
        kk.append([int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k])
    return bounds, kk


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    bounds, kk = _coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx, (xmin, n) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[xmin + x] * kk[xx][x]
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def pil_bicubic_resize_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """ImagingResample for 8-bit images: horizontal pass, then vertical pass, uint8 between them; a pass whose size does
    not change is skipped.  img (H, W, C) uint8."""
    if img.shape[1] != out_w:
        img = _pass(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, 0)
    return img


def resize_output_size(height: int, width: int, shortest_edge: int):
    """transformers 4.x get_resize_output_image_size(size=int, default_to_square=False) -> (height, width)."""
    short, long = (width, height) if width <= height else (height, width)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if width <= height else (new_short, new_long)


def bit_preprocess_pil(images, shortest_edge=256, crop=(224, 224), rescale=1 / 255.0,
                       mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), return_u8=False):
    """transformers 4.x BitImageProcessor.preprocess restated on top of Pillow itself -> (T, 3, h, w) float32."""
    from PIL import Image

    outs, u8s = [], []
    for im in images:
        im = im.convert("RGB")                                                      # do_convert_rgb
        a = np.asarray(im, dtype=np.uint8)
        oh, ow = resize_output_size(a.shape[0], a.shape[1], shortest_edge)
        r = np.asarray(Image.fromarray(a, "RGB").resize((ow, oh), resample=Image.BICUBIC))   # image_transforms.resize
        top, left = (oh - crop[0]) // 2, (ow - crop[1]) // 2                        # image_transforms.center_crop
        c = r[top:top + crop[0], left:left + crop[1]]
        x = (c.astype(np.float64) * rescale).astype(np.float32)                     # image_transforms.rescale
        x = (x - np.array(mean, dtype=np.float32)) / np.array(std, dtype=np.float32)  # image_transforms.normalize
        outs.append(np.transpose(x, (2, 0, 1)))
        u8s.append(c)
    pv = np.stack(outs).astype(np.float32)
    return (pv, np.stack(u8s)) if return_u8 else pv


# ---- ImagePreprocessor (actionmesh/preprocessing/image_processor.py:26-146) restated with numpy only ------------------------
def frame_preprocess(frames_rgba, independent_cropping: bool = False, padding_ratio: float = 0.1):
    """frames_rgba: list of (H, W, 4) uint8 arrays -> list of (H', W', 3) uint8 arrays, following load_image (:26-66),
    aggregate_bboxes (:69-77), apply_padding (:80-101) and the uint8 conversion of process_images (:142-146)."""
    import numpy as np

    bg = np.array([1.0, 1.0, 1.0]).astype(np.float32)
    comps, boxes = [], []
    for img in frames_rgba:
        rgb, alpha = img[..., :3], img[..., 3]
        total = alpha.size
        min_count = int(total * 0.01)
        fg = np.count_nonzero(alpha > 127)
        if not (total - fg >= min_count and fg >= min_count):
            raise ValueError("Invalid alpha channel: insufficient foreground/background")
        a = (alpha.astype(np.float32) * (1.0 / 255.0))[..., None]
        comps.append(rgb.astype(np.float32) * (1.0 / 255.0) * a + bg * (1.0 - a))
        mask = alpha > 0
        rows, cols = np.nonzero(mask.any(axis=1))[0], np.nonzero(mask.any(axis=0))[0]
        boxes.append((cols[0], rows[0], cols[-1] - cols[0] + 1, rows[-1] - rows[0] + 1))
    if not independent_cropping:
        x0, y0 = min(b[0] for b in boxes), min(b[1] for b in boxes)
        x1, y1 = max(b[0] + b[2] for b in boxes), max(b[1] + b[3] for b in boxes)
        boxes = [(x0, y0, x1 - x0, y1 - y0)] * len(boxes)
    out = []
    for comp, (x, y, w, h) in zip(comps, boxes):
        crop = comp[y:y + h, x:x + w]
        m = max(w, h)
        pb = int(m * padding_ratio)
        px, py = pb + (m - w) // 2, pb + (m - h) // 2
        padded = np.pad(crop, ((py, py), (px, px), (0, 0)), mode="constant", constant_values=1.0)
        out.append((padded * np.float32(255)).astype(np.uint8))
    return out
