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
from typing import Dict, List, Optional

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


def text_forward(arch: TextArch, w: Dict[str, torch.Tensor], tokens: torch.Tensor, normalize: bool = True,
                 lora: Optional[Dict[int, List[torch.Tensor]]] = None, lora_scale: float = 2.0) -> torch.Tensor:
    """tokens int64 [n, context] -> class features [n, proj] (L2-normalised like custom_clip.py:662).
    `lora` {layer: (A_q, B_q, A_v, B_v)}: peft LoRA on q_proj / v_proj of the text tower (`--lora_encoder text`,
    clip/custom_clip.py:602-606): y = W h + b + lora_scale * B (A h); lora_scale = alpha / r = 32 / 16."""
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
        if lora is not None and i in lora:
            a_q, b_q, a_v, b_v = lora[i]
            q = q + (lora_scale * F.linear(F.linear(h, a_q), b_q)).view(n, T, H, dh).transpose(1, 2)
            v = v + (lora_scale * F.linear(F.linear(h, a_v), b_v)).view(n, T, H, dh).transpose(1, 2)
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


def text_lora_init(arch: TextArch, layers: range, rank: int = 16, seed: int = 0) -> Dict[int, List[torch.Tensor]]:
    """LoRA_AB(text_encoder, init_method='xavier') (clip/custom_clip.py:152-200): A ~ xavier_normal_, B = 0."""
    g = torch.Generator().manual_seed(seed)
    std = math.sqrt(2.0 / (arch.width + rank))
    out = {}
    for i in layers:
        a_q = torch.randn(rank, arch.width, generator=g) * std
        a_v = torch.randn(rank, arch.width, generator=g) * std
        out[i] = [a_q, torch.zeros(arch.width, rank), a_v, torch.zeros(arch.width, rank)]
    return out


def adapt_and_predict_text_lora(varch, vw: Dict[str, torch.Tensor], tarch: TextArch, tw: Dict[str, torch.Tensor],
                                tokens: torch.Tensor, images: torch.Tensor, logit_scale: float,
                                lora0: Dict[int, List[torch.Tensor]], lora_scale: float = 2.0, head: str = "tpt",
                                tta_steps: int = 1, selection_p: float = 0.1, lr: float = 5e-3, margin_e0: float = 0.4,
                                forced_idx: Optional[torch.Tensor] = None):
    """One test sample with the adapter on the TEXT tower (`--lora_encoder text`): the image features of the views are
    frozen (clip/custom_clip.py:672-673, no_grad), the class features are recomputed WITH gradient in every forward
    (:677-678), the loss heads, AdamW and the reset are those of the image route (ttl.py:70-110, 338-352).  Returns the
    AdaptResult of oracle.ttl_oracle (idx / grads / lora refer to the text-tower factors)."""
    from oracle import ttl_oracle as O
    lora = {i: [t.clone().requires_grad_(True) for t in ts] for i, ts in lora0.items()}
    st = O.AdamWState()
    with torch.no_grad():
        feats = O.vision_forward(varch, vw, images, None, 0.0)          # plain image tower: peft wraps the text tower only
    n_opt_steps = tta_steps * tta_steps if head == "deyo" else tta_steps
    sel, first_logits, ent0, losses, grads = forced_idx, None, None, [], {}
    for _ in range(n_opt_steps):
        text = text_forward(tarch, tw, tokens, True, lora, lora_scale)
        logits = O.clip_logits(feats, text, logit_scale)
        if first_logits is None:
            first_logits = logits.detach().clone()
            ent0 = O.softmax_entropy(first_logits)
        if head == "tpt":
            if sel is None:
                _, sel = O.select_confident_samples(logits.detach(), selection_p)
            loss = O.avg_entropy(logits[sel].float())
        else:
            loss = O.deyo_loss(logits, margin_e0)
            if sel is None:
                sel = torch.where(ent0 <= math.log(1000))[0]
        flat = [t for ts in lora.values() for t in ts]
        gs = torch.autograd.grad(loss, flat, allow_unused=True)
        it = iter(gs)
        grads = {i: [next(it) for _ in ts] for i, ts in lora.items()}
        grads = {i: [torch.zeros_like(p) if g is None else g for g, p in zip(gl, lora[i])] for i, gl in grads.items()}
        with torch.no_grad():
            O.adamw_step({i: [t for t in ts] for i, ts in lora.items()}, grads, st, lr=lr)
        losses.append(float(loss.detach()))
    with torch.no_grad():
        lora_d = {i: [t.detach() for t in ts] for i, ts in lora.items()}
        text = text_forward(tarch, tw, tokens, True, lora_d, lora_scale)
        pred = O.clip_logits(feats[:1], text, logit_scale)
    return O.AdaptResult(first_logits, ent0, sel, losses[-1] if losses else float("nan"), grads, lora_d, pred, losses)
