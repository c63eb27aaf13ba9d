"""Golden fixture pinning the OpenAI-format weight mapping (ttl_b200/weights.py) and the oracle's
encoder math to the reference's OWN OpenAI-format model: clip/model.py `VisionTransformer` (nn.MultiheadAttention,
QuickGELU, `x @ proj`) is instantiated at a tiny geometry with seeded random weights and run on seeded images; the state
dict, the images and the output features go into tests/golden/openai_vit_tiny.npz.
Dev-container only (imports /root/reference/clip/model.py behind oracle/ref_shim.py).  Usage: python oracle/make_golden_openai_format.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402

GEOM = dict(input_resolution=32, patch_size=16, width=64, layers=2, heads=1, output_dim=24)


def main() -> None:
    R.install(None)
    from clip.model import VisionTransformer     # the reference's module
    torch.manual_seed(4321)
    vit = VisionTransformer(**GEOM).float().eval()
    with torch.no_grad():
        for n, p in vit.named_parameters():       # biases / LN affine away from their 0/1 defaults so every slot matters
            if n.endswith("bias") or "ln_" in n:
                p.add_(0.1 * torch.randn_like(p))
    imgs = torch.randn(3, 3, 32, 32)
    with torch.no_grad():
        feats = vit(imgs)
    rec = {"sd::visual." + k: v.detach().numpy() for k, v in vit.state_dict().items()}
    rec.update(images=imgs.numpy(), features=feats.numpy(), **{f"geom_{k}": np.int64(v) for k, v in GEOM.items()})
    path = os.path.join(ROOT, "tests", "golden", "openai_vit_tiny.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", "features", tuple(feats.shape))


if __name__ == "__main__":
    main()
