"""CPU oracle of the class-feature builder ("next" row N2): fp32 PyTorch restatement of the text tower the
reference calls through HF `CLIPModel.get_text_features` (clip/custom_clip.py:73-82, 651-663; transformers
models/clip/modeling_clip.py CLIPTextEmbeddings / CLIPEncoderLayer / CLIPTextTransformer: causal mask, pooling at the
EOT position = argmax of the token ids, text_projection without bias) followed by the L2 normalisation of :662.

TEST INFRASTRUCTURE ONLY (imported by tests/ only).  Parity pin: tests/test_text_oracle.py compares it with
transformers' own CLIPTextModelWithProjection (the un-vendored third-party code the reference runs; 5.5.0 in this image) on
seeded random weights, and the tokenizer restatement with the reference's SimpleTokenizer where /root/reference exists."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class TextArch:
    vocab: int = 49408
    context: int = 77
    width: int = 512
    layers: int = 12
    heads: int = 8
    mlp: int = 2048
    proj: int = 512
    ln_eps: float = 1e-5


TEXT_ARCHS = {"ViT-B/16": TextArch(), "ViT-L/14": TextArch(width=768, heads=12, mlp=3072, proj=768),
              "tiny": TextArch(vocab=1000, context=16, width=128, layers=2, heads=2, mlp=512, proj=64)}


def make_synthetic_text_weights(arch: TextArch, seed: int = 4321) -> Dict[str, torch.Tensor]:
    """HF-style init (modeling_clip.py _init_weights) with small random biases / LN affine so every slot is exercised."""
    g = torch.Generator().manual_seed(seed)
    d, L = arch.width, arch.layers
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    w = {"text_model.embeddings.token_embedding.weight": rn(arch.vocab, d, std=0.02),
         "text_model.embeddings.position_embedding.weight": rn(arch.context, d, std=0.02),
         "text_model.final_layer_norm.weight": 1 + rn(d, std=0.05), "text_model.final_layer_norm.bias": rn(d, std=0.05),
         "text_projection.weight": rn(arch.proj, d, std=d ** -0.5)}
    in_std, out_std, fc_std = d ** -0.5 * (2 * L) ** -0.5, d ** -0.5, (2 * d) ** -0.5
    for i in range(L):
        p = f"text_model.encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj"):
            w[p + f"self_attn.{nm}.weight"], w[p + f"self_attn.{nm}.bias"] = rn(d, d, std=in_std), rn(d, std=0.02)
        w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"] = rn(d, d, std=out_std), rn(d, std=0.02)
        w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"] = rn(arch.mlp, d, std=fc_std), rn(arch.mlp, std=0.02)
        w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"] = rn(d, arch.mlp, std=in_std), rn(d, std=0.02)
        for ln in ("layer_norm1", "layer_norm2"):
            w[p + ln + ".weight"], w[p + ln + ".bias"] = 1 + rn(d, std=0.05), rn(d, std=0.05)
    return w


def make_synthetic_tokens(n: int, arch: TextArch, seed: int = 9) -> torch.Tensor:
    """clip.tokenize layout (clip/clip.py:196-232): SOT, 2..context-2 ids, EOT (= highest id of the vocabulary), zeros."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(n, arch.context, dtype=torch.int64)
    sot, eot = arch.vocab - 2, arch.vocab - 1
    for i in range(n):
        k = int(torch.randint(2, arch.context - 1, (1,), generator=g))
        out[i, 0] = sot
        out[i, 1:1 + k - 1] = torch.randint(1, arch.vocab - 2, (k - 1,), generator=g)
        out[i, k] = eot
    return out


def text_forward(arch: TextArch, w: Dict[str, torch.Tensor], tokens: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """tokens int64 [n, context] -> class features [n, proj] (L2-normalised like custom_clip.py:662)."""
    n, T = tokens.shape
    d, H = arch.width, arch.heads
    dh = d // H
    x = w["text_model.embeddings.token_embedding.weight"][tokens] + w["text_model.embeddings.position_embedding.weight"][:T]
    mask = torch.full((T, T), float("-inf")).triu(1)
    for i in range(arch.layers):
        p = f"text_model.encoder.layers.{i}."
        h = F.layer_norm(x, (d,), w[p + "layer_norm1.weight"], w[p + "layer_norm1.bias"], arch.ln_eps)
        q = F.linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"]).view(n, T, H, dh).transpose(1, 2)
        k = F.linear(h, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"]).view(n, T, H, dh).transpose(1, 2)
        v = F.linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"]).view(n, T, H, dh).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5 + mask, dim=-1)
        o = (att @ v).transpose(1, 2).reshape(n, T, d)
        x = x + F.linear(o, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (d,), w[p + "layer_norm2.weight"], w[p + "layer_norm2.bias"], arch.ln_eps)
        h = F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])
        x = x + F.linear(h * torch.sigmoid(1.702 * h), w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
    x = F.layer_norm(x, (d,), w["text_model.final_layer_norm.weight"], w["text_model.final_layer_norm.bias"], arch.ln_eps)
    pooled = x[torch.arange(n), tokens.argmax(dim=-1)]
    feats = F.linear(pooled, w["text_projection.weight"])
    return feats / feats.norm(dim=-1, keepdim=True) if normalize else feats
