"""CPU oracle for the view generator ("next" row N1): the reference builds the 64 views of a test image on
the CPU with PIL + torchvision (data/datautils.py:98-157, ttl.py:232-241):

    view 0      = Normalize(ToTensor(CenterCrop(224)(Resize(224, BICUBIC)(img))))
    views 1..63 = Normalize(ToTensor(RandomHorizontalFlip()(RandomResizedCrop(224)(img))))     (AugMix op list is empty)

TEST INFRASTRUCTURE ONLY (imported by tests/ only).  This file restates Pillow's antialiased two-pass resampler for 8-bit
images (Pillow src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
Vertical_8bpc; pinned Pillow 12.2 in this image) in numpy integer arithmetic.  Parity pin: tests/test_views_oracle.py checks
it bit-exactly against Pillow itself (Image.resize on random crops and sizes) and against the torchvision transforms the
reference composes."""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
BILINEAR, BICUBIC = 0, 1
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _filter(mode: int, x: float) -> float:
    if x < 0.0:
        x = -x
    if mode == BILINEAR:
        return 1.0 - x if x < 1.0 else 0.0
    a = -0.5
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int, mode: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """bounds [out,2] (xmin, count) and fixed-point coefficients [out, ksize] (int32), as Pillow computes them for a resize
    of a whole image axis (box = (0, in_size))."""
    support0 = 1.0 if mode == BILINEAR else 2.0
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_filter(mode, (x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_u8(img: np.ndarray, out_h: int, out_w: int, mode: int) -> np.ndarray:
    """img uint8 [H,W,C] -> uint8 [out_h,out_w,C]: horizontal pass to uint8, then vertical pass (Pillow's order)."""
    H, W, C = img.shape
    bh, kh, _ = precompute_coeffs(W, out_w, mode)
    bv, kv, _ = precompute_coeffs(H, out_h, mode)
    tmp = np.empty((H, out_w, C), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_w):
        x0, n = bh[xx]
        acc = (src[:, x0:x0 + n, :] * kh[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        tmp[:, xx, :] = _clip8(acc)
    out = np.empty((out_h, out_w, C), dtype=np.uint8)
    t64 = tmp.astype(np.int64)
    for yy in range(out_h):
        y0, n = bv[yy]
        acc = (t64[y0:y0 + n] * kv[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
        out[yy] = _clip8(acc)
    return out


def to_normalized(u8: np.ndarray, mean: Sequence[float] = CLIP_MEAN, std: Sequence[float] = CLIP_STD) -> np.ndarray:
    """ToTensor + Normalize in fp32, in torchvision's operation order: (x / 255 - mean) / std -> [C,H,W]."""
    x = u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return (x - m) / s


def resized_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision _compute_resized_output_size for an int size (smaller edge -> size)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def clean_view(img: np.ndarray, size: int = 224) -> np.ndarray:
    """ttl.py:232-241 base transform: Resize(size, BICUBIC) + CenterCrop(size) -> uint8 [size,size,3]."""
    H, W, _ = img.shape
    nh, nw = resized_size(H, W, size)
    r = resize_u8(img, nh, nw, BICUBIC)
    top, left = int(round((nh - size) / 2.0)), int(round((nw - size) / 2.0))
    return r[top:top + size, left:left + size]


def crop_view(img: np.ndarray, i: int, j: int, h: int, w: int, flip: bool, size: int = 224) -> np.ndarray:
    """RandomResizedCrop(size) with the drawn box (i, j, h, w) + optional horizontal flip -> uint8 [size,size,3]."""
    r = resize_u8(img[i:i + h, j:j + w], size, size, BILINEAR)
    return r[:, ::-1] if flip else r


def make_views(img: np.ndarray, boxes: Sequence[Tuple[int, int, int, int, int]], size: int = 224) -> np.ndarray:
    """[1 + len(boxes), 3, size, size] fp32: the clean view followed by the crop views (boxes: i, j, h, w, flip)."""
    views = [to_normalized(clean_view(img, size))]
    for (i, j, h, w, f) in boxes:
        views.append(to_normalized(crop_view(img, i, j, h, w, bool(f), size)))
    return np.stack(views)
