"""Host side of the GPU view generator: the part of the reference's AugMixAugmenter that must stay on the host.

The reference's DataLoader workers turn one PIL image into 64 normalised fp32 views (data/datautils.py:98-157,
ttl.py:232-241).  Here the host only *draws* the views -- the RandomResizedCrop box and the flip coin, with torchvision's
own sampler so the torch RNG stream is consumed exactly as the reference's transforms consume it -- and hands the decoded
uint8 image plus an int32 spec table to libttl_b200, which resamples on the device bit-exactly as Pillow would.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

SPEC_FIELDS = 6   # kind, top, left, height, width, flip  (struct ttl_view_spec)


class ViewSpecSampler:
    """Counterpart of `AugMixAugmenter(base_transform, preprocess, n_views, augmix)` (data/datautils.py:129-157) whose
    __call__ returns (uint8 image [H,W,3], int32 specs [1 + n_views, 6]) instead of the list of view tensors: row 0 is the
    clean view (base_transform = Resize(BICUBIC) + CenterCrop), rows 1.. are get_preaugment()'s RandomResizedCrop +
    RandomHorizontalFlip draws.  The augmentation list is empty in the reference (datautils.py:135-138), so nothing else
    is sampled."""

    def __init__(self, n_views: int = 63, scale: Tuple[float, float] = (0.08, 1.0),
                 ratio: Tuple[float, float] = (3.0 / 4.0, 4.0 / 3.0), flip_p: float = 0.5):
        self.n_views, self.scale, self.ratio, self.flip_p = n_views, list(scale), list(ratio), flip_p

    def __call__(self, img):
        from torchvision.transforms import RandomResizedCrop
        arr = np.asarray(img.convert("RGB") if hasattr(img, "convert") else img, dtype=np.uint8)
        if arr.ndim != 3 or arr.shape[2] != 3:
            raise ValueError("expected an RGB image [H,W,3]")
        if not arr.flags.writeable:      # a PIL image exposes a read-only buffer; torch.from_numpy wants a writable one
            arr = arr.copy()
        h, w = int(arr.shape[0]), int(arr.shape[1])
        probe = torch.empty(3, h, w, dtype=torch.uint8, device="meta")   # get_params only reads the size
        specs = np.zeros((1 + self.n_views, SPEC_FIELDS), dtype=np.int32)
        specs[0] = (L.VIEW_CLEAN, 0, 0, h, w, 0)
        for v in range(self.n_views):
            i, j, ch, cw = RandomResizedCrop.get_params(probe, self.scale, self.ratio)   # RandomResizedCrop.forward
            flip = int(torch.rand(1) < self.flip_p)                                       # RandomHorizontalFlip.forward
            specs[1 + v] = (L.VIEW_CROP, i, j, ch, cw, flip)
        return np.ascontiguousarray(arr), specs


def pack_specs(specs: Sequence[np.ndarray]) -> np.ndarray:
    """[n_images][n_views, 6] -> one contiguous int32 [n_images, n_views, 6] table (struct ttl_view_spec array)."""
    out = np.ascontiguousarray(np.stack([np.asarray(s, dtype=np.int32) for s in specs]))
    if out.ndim != 3 or out.shape[2] != SPEC_FIELDS:
        raise ValueError("specs must be [n_views, 6] per image")
    return out
