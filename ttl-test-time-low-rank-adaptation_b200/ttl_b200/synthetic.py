"""Seeded synthetic inputs for benchmarks and smoke runs (there are no checkpoints or datasets offline): random-init
CLIP ViT weights with HF `_init_weights` standard deviations, Xavier-normal LoRA A / zero B (clip/custom_clip.py:
139-200), unit-norm text features.  Same seeds give the same tensors as the test oracle's generators."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch

from .engine import ARCH_GEOMETRY


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def synthetic_vit_weights(arch: str = "ViT-B/16", seed: int = 1234, affine_noise: float = 0.05) -> Dict[str, torch.Tensor]:
    geo = ARCH_GEOMETRY[arch]
    d, L, F, P, p = geo["width"], geo["layers"], geo["mlp_dim"], geo["proj_dim"], geo["patch"]
    tokens = (geo["image_size"] // p) ** 2 + 1
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    pre = "vision_model."
    w[pre + "embeddings.class_embedding"] = _randn(g, d, std=d ** -0.5)
    w[pre + "embeddings.patch_embedding.weight"] = _randn(g, d, 3, p, p, std=0.02)
    w[pre + "embeddings.position_embedding.weight"] = _randn(g, tokens, d, std=0.02)

    def ln(name):
        w[name + ".weight"] = 1.0 + _randn(g, d, std=affine_noise)
        w[name + ".bias"] = _randn(g, d, std=affine_noise)

    ln(pre + "pre_layrnorm")
    in_std, out_std, fc_std = d ** -0.5 * (2 * L) ** -0.5, d ** -0.5, (2 * d) ** -0.5
    for i in range(L):
        q = f"{pre}encoder.layers.{i}."
        ln(q + "layer_norm1")
        for nm in ("q_proj", "k_proj", "v_proj"):
            w[q + f"self_attn.{nm}.weight"] = _randn(g, d, d, std=in_std)
            w[q + f"self_attn.{nm}.bias"] = _randn(g, d, std=affine_noise * 0.2)
        w[q + "self_attn.out_proj.weight"] = _randn(g, d, d, std=out_std)
        w[q + "self_attn.out_proj.bias"] = _randn(g, d, std=affine_noise * 0.2)
        ln(q + "layer_norm2")
        w[q + "mlp.fc1.weight"] = _randn(g, F, d, std=fc_std)
        w[q + "mlp.fc1.bias"] = _randn(g, F, std=affine_noise * 0.2)
        w[q + "mlp.fc2.weight"] = _randn(g, d, F, std=in_std)
        w[q + "mlp.fc2.bias"] = _randn(g, d, std=affine_noise * 0.2)
    ln(pre + "post_layernorm")
    w["visual_projection.weight"] = _randn(g, P, d, std=d ** -0.5)
    return w


def synthetic_lora_init(arch: str = "ViT-B/16", rank: int = 16, layers: Sequence[int] = (9, 11), seed: int = 0
                        ) -> Dict[int, List[torch.Tensor]]:
    d = ARCH_GEOMETRY[arch]["width"]
    g = torch.Generator().manual_seed(seed)
    std = math.sqrt(2.0 / (d + rank))
    out = {}
    for i in range(layers[0], layers[1] + 1):
        a_q = _randn(g, rank, d, std=std)
        a_v = _randn(g, rank, d, std=std)
        out[i] = [a_q, torch.zeros(d, rank), a_v, torch.zeros(d, rank)]
    return out


def synthetic_text_features(n_classes: int, proj: int, seed: int = 11) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    t = _randn(g, n_classes, proj)
    return t / t.norm(dim=-1, keepdim=True)


def synthetic_text_weights(arch: str = "ViT-B/16", seed: int = 4321, affine_noise: float = 0.05) -> Dict[str, torch.Tensor]:
    """Seeded random-init CLIP text tower with HF `_init_weights` standard deviations (HF names); for `--lora_encoder text`
    runs without a checkpoint (--synthetic / --random_init)."""
    from .engine import TEXT_TOWER_GEOMETRY
    geo = TEXT_TOWER_GEOMETRY[arch]
    d, L, F, P, ctx, vocab = geo["width"], geo["layers"], geo["mlp_dim"], geo["proj_dim"], geo["context"], geo["vocab"]
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    pre = "text_model."
    w[pre + "embeddings.token_embedding.weight"] = _randn(g, vocab, d, std=0.02)
    w[pre + "embeddings.position_embedding.weight"] = _randn(g, ctx, d, std=0.02)

    def ln(name):
        w[name + ".weight"] = 1.0 + _randn(g, d, std=affine_noise)
        w[name + ".bias"] = _randn(g, d, std=affine_noise)

    in_std, out_std, fc_std = d ** -0.5 * (2 * L) ** -0.5, d ** -0.5, (2 * d) ** -0.5
    for i in range(L):
        q = f"{pre}encoder.layers.{i}."
        ln(q + "layer_norm1")
        for nm in ("q_proj", "k_proj", "v_proj"):
            w[q + f"self_attn.{nm}.weight"] = _randn(g, d, d, std=in_std)
            w[q + f"self_attn.{nm}.bias"] = _randn(g, d, std=affine_noise * 0.2)
        w[q + "self_attn.out_proj.weight"] = _randn(g, d, d, std=out_std)
        w[q + "self_attn.out_proj.bias"] = _randn(g, d, std=affine_noise * 0.2)
        ln(q + "layer_norm2")
        w[q + "mlp.fc1.weight"] = _randn(g, F, d, std=fc_std)
        w[q + "mlp.fc1.bias"] = _randn(g, F, std=affine_noise * 0.2)
        w[q + "mlp.fc2.weight"] = _randn(g, d, F, std=in_std)
        w[q + "mlp.fc2.bias"] = _randn(g, d, std=affine_noise * 0.2)
    ln(pre + "final_layer_norm")
    w["text_projection.weight"] = _randn(g, P, d, std=d ** -0.5)
    return w


class HashTokenizer:
    """Stand-in for the CLIP BPE table in random-init runs (the merge table is data of the reference, not shipped): stable ids
    from the words of a prompt, clip.tokenize layout (SOT, ids, EOT = highest id, zero padding)."""

    def __init__(self, vocab: int = 49408, context: int = 77):
        self.vocab, self.context = vocab, context

    def __call__(self, prompts):
        import zlib
        out = torch.zeros(len(prompts), self.context, dtype=torch.long)
        for i, p in enumerate(prompts):
            ids = [1 + zlib.crc32(wd.encode()) % (self.vocab - 3) for wd in p.replace(".", " .").split()][: self.context - 2]
            row = [self.vocab - 2] + ids + [self.vocab - 1]
            out[i, :len(row)] = torch.tensor(row)
        return out
