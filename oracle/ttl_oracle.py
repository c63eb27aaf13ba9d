"""CPU oracle for the TTL per-sample test-time adaptation path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may.  It is a plain PyTorch fp32 restatement (CPU, autograd for the
backward) of the reference algorithm, self-contained so it travels to the GPU box where
``/root/reference`` does not exist.

Parity pin: the reference ships no tests or golden vectors for this path (SURVEY.md §4,
§8c) and its arithmetic lives in un-vendored third-party code (HF ``transformers`` CLIP,
``peft`` LoRA, ``torch.optim.AdamW``).  The oracle is therefore pinned against OUTPUTS OF
THE REFERENCE ITSELF, run unmodified in the dev container behind import shims
(``oracle/ref_shim.py``; generator ``oracle/make_golden.py``; fixtures
``tests/golden/*.npz``; check ``tests/test_oracle_golden.py``).

What each function follows (paths relative to /root/reference):
  * ``vision_forward``            HF ``modeling_clip.py`` CLIPVisionEmbeddings.forward,
                                  CLIPAttention.forward (+eager_attention_forward),
                                  CLIPMLP.forward, CLIPEncoderLayer.forward,
                                  CLIPVisionTransformer.forward, CLIPModel.get_image_features;
                                  LoRA branch: peft ``Linear.forward``
                                  (``base(x) + lora_B(lora_A(x)) * alpha/r``) as configured at
                                  clip/custom_clip.py:583-591.
  * ``clip_logits``               clip/custom_clip.py:665-694 (``inference``).
  * ``select_confident_samples``  ttl.py:50-54.
  * ``avg_entropy``               ttl.py:56-61.
  * ``softmax_entropy`` / ``deyo_loss``   deyo.py:85-90, 93-110, 159-181 (default flags).
  * ``adamw_step``                torch.optim.AdamW single-tensor rule with the defaults used at
                                  ttl.py:218 (betas .9/.999, eps 1e-8, wd 1e-2).
  * ``lora_init``                 clip/custom_clip.py:139-200 (Xavier-normal A, B = 0).
  * ``adapt_and_predict``         ttl.py:338-352 + ttl.py:70-110.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- architecture
@dataclass(frozen=True)
class VitArch:
    name: str = "ViT-B/16"
    image_size: int = 224
    patch: int = 16
    width: int = 768
    layers: int = 12
    heads: int = 12
    mlp: int = 3072
    proj: int = 512
    ln_eps: float = 1e-5

    @property
    def tokens(self) -> int:
        return (self.image_size // self.patch) ** 2 + 1

    @property
    def head_dim(self) -> int:
        return self.width // self.heads


ARCHS = {
    "ViT-B/16": VitArch(),
    "ViT-L/14": VitArch("ViT-L/14", 224, 14, 1024, 24, 16, 4096, 768),
    # tiny configuration for fast CPU tests; same code path, nothing special-cased
    "ViT-tiny": VitArch("ViT-tiny", 64, 16, 128, 4, 2, 512, 64),
}


@dataclass
class LoraSpec:
    rank: int = 16
    alpha: float = 32.0
    layer_lo: int = 9
    layer_hi: int = 11  # inclusive (ttl.py:402 "inclusive range")

    @property
    def scale(self) -> float:
        return self.alpha / self.rank

    def layers(self) -> range:
        return range(self.layer_lo, self.layer_hi + 1)


# ----------------------------------------------------------------------------- synthetic weights
def _randn(gen: torch.Generator, *shape: int, std: float = 1.0) -> torch.Tensor:
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std


def make_synthetic_weights(arch: VitArch, seed: int = 1234, affine_noise: float = 0.05) -> Dict[str, torch.Tensor]:
    """Seeded random-init CLIP vision tower with HF ``_init_weights`` standard deviations
    (HF modeling_clip.py ``CLIPPreTrainedModel._init_weights``), except that biases and LayerNorm
    affine parameters get small random values so those code paths are exercised
    (SURVEY.md §8d "Synthetic inputs").  Keys are HF ``state_dict`` names."""
    g = torch.Generator().manual_seed(seed)
    d, L, F_, P = arch.width, arch.layers, arch.mlp, arch.proj
    w: Dict[str, torch.Tensor] = {}
    pre = "vision_model."
    w[pre + "embeddings.class_embedding"] = _randn(g, d, std=d ** -0.5)
    w[pre + "embeddings.patch_embedding.weight"] = _randn(g, d, 3, arch.patch, arch.patch, std=0.02)
    w[pre + "embeddings.position_embedding.weight"] = _randn(g, arch.tokens, d, std=0.02)

    def ln(name: str) -> None:
        w[name + ".weight"] = 1.0 + _randn(g, d, std=affine_noise)
        w[name + ".bias"] = _randn(g, d, std=affine_noise)

    ln(pre + "pre_layrnorm")
    in_std = d ** -0.5 * (2 * L) ** -0.5
    out_std = d ** -0.5
    fc_std = (2 * d) ** -0.5
    for i in range(L):
        p = f"{pre}encoder.layers.{i}."
        ln(p + "layer_norm1")
        for nm in ("q_proj", "k_proj", "v_proj"):
            w[p + f"self_attn.{nm}.weight"] = _randn(g, d, d, std=in_std)
            w[p + f"self_attn.{nm}.bias"] = _randn(g, d, std=affine_noise * 0.2)
        w[p + "self_attn.out_proj.weight"] = _randn(g, d, d, std=out_std)
        w[p + "self_attn.out_proj.bias"] = _randn(g, d, std=affine_noise * 0.2)
        ln(p + "layer_norm2")
        w[p + "mlp.fc1.weight"] = _randn(g, F_, d, std=fc_std)
        w[p + "mlp.fc1.bias"] = _randn(g, F_, std=affine_noise * 0.2)
        w[p + "mlp.fc2.weight"] = _randn(g, d, F_, std=in_std)
        w[p + "mlp.fc2.bias"] = _randn(g, d, std=affine_noise * 0.2)
    ln(pre + "post_layernorm")
    w["visual_projection.weight"] = _randn(g, P, d, std=d ** -0.5)
    return w


def lora_init(arch: VitArch, spec: LoraSpec, seed: int = 0) -> Dict[int, List[torch.Tensor]]:
    """LoRA factors for every trainable layer as ``[A_q, B_q, A_v, B_v]`` (the tuple order of
    clip/custom_clip.py:193-200).  A ~ xavier_normal_ (std = sqrt(2/(fan_in+fan_out))), B = 0."""
    g = torch.Generator().manual_seed(seed)
    d, r = arch.width, spec.rank
    std = math.sqrt(2.0 / (d + r))
    out: Dict[int, List[torch.Tensor]] = {}
    for i in spec.layers():
        a_q = _randn(g, r, d, std=std)
        a_v = _randn(g, r, d, std=std)
        out[i] = [a_q, torch.zeros(d, r), a_v, torch.zeros(d, r)]
    return out


def make_synthetic_views(n_views: int, image_size: int = 224, seed: int = 7) -> torch.Tensor:
    """One smooth base image (roughly unit variance, like CLIP-normalised pixels); view 0 is the
    centre crop, views 1.. are random-resized-crop + horizontal flip of it, mirroring what
    data/datautils.py:98-108,129-157 feeds the loop (its AugMix op list is empty)."""
    g = torch.Generator().manual_seed(seed)
    big = int(image_size * 1.25)
    lo = F.interpolate(_randn(g, 1, 3, 7, 7), size=(big, big), mode="bicubic", align_corners=False)
    hi = F.interpolate(_randn(g, 1, 3, 56, 56), size=(big, big), mode="bilinear", align_corners=False)
    base = (lo + 0.5 * hi)[0]
    views = []
    off = (big - image_size) // 2
    views.append(base[:, off:off + image_size, off:off + image_size])
    area = big * big
    for _ in range(n_views - 1):
        for _try in range(10):
            s = float(torch.empty(1).uniform_(0.08, 1.0, generator=g)) * area
            logr = float(torch.empty(1).uniform_(math.log(3 / 4), math.log(4 / 3), generator=g))
            ar = math.exp(logr)
            cw, ch = int(round(math.sqrt(s * ar))), int(round(math.sqrt(s / ar)))
            if 0 < cw <= big and 0 < ch <= big:
                break
        else:
            cw = ch = big
        top = int(torch.randint(0, big - ch + 1, (1,), generator=g))
        left = int(torch.randint(0, big - cw + 1, (1,), generator=g))
        crop = base[:, top:top + ch, left:left + cw][None]
        v = F.interpolate(crop, size=(image_size, image_size), mode="bilinear", align_corners=False)[0]
        if float(torch.rand(1, generator=g)) < 0.5:
            v = v.flip(-1)
        views.append(v)
    return torch.stack(views).contiguous()


# ----------------------------------------------------------------------------- model
def _lora_linear(x, w, b, lora_a, lora_b, scale):
    y = F.linear(x, w, b)
    if lora_a is not None:
        y = y + F.linear(F.linear(x, lora_a), lora_b) * scale
    return y


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


def vision_embed(arch: VitArch, w: Dict[str, torch.Tensor], images: torch.Tensor) -> torch.Tensor:
    pre = "vision_model."
    x = F.conv2d(images, w[pre + "embeddings.patch_embedding.weight"], stride=arch.patch)
    x = x.flatten(2).transpose(1, 2)
    cls = w[pre + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + w[pre + "embeddings.position_embedding.weight"]
    return F.layer_norm(x, (arch.width,), w[pre + "pre_layrnorm.weight"], w[pre + "pre_layrnorm.bias"], arch.ln_eps)


def encoder_layer(arch: VitArch, w: Dict[str, torch.Tensor], i: int, x: torch.Tensor,
                  lora: Optional[Sequence[torch.Tensor]], scale: float) -> torch.Tensor:
    p = f"vision_model.encoder.layers.{i}."
    B, N, d = x.shape
    H, dh = arch.heads, arch.head_dim
    h = F.layer_norm(x, (d,), w[p + "layer_norm1.weight"], w[p + "layer_norm1.bias"], arch.ln_eps)
    a_q, b_q, a_v, b_v = lora if lora is not None else (None, None, None, None)
    q = _lora_linear(h, w[p + "self_attn.q_proj.weight"], w[p + "self_attn.q_proj.bias"], a_q, b_q, scale)
    k = F.linear(h, w[p + "self_attn.k_proj.weight"], w[p + "self_attn.k_proj.bias"])
    v = _lora_linear(h, w[p + "self_attn.v_proj.weight"], w[p + "self_attn.v_proj.bias"], a_v, b_v, scale)
    q = q.view(B, N, H, dh).transpose(1, 2)
    k = k.view(B, N, H, dh).transpose(1, 2)
    v = v.view(B, N, H, dh).transpose(1, 2)
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5), dim=-1)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, N, d)
    x = x + F.linear(o, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
    h = F.layer_norm(x, (d,), w[p + "layer_norm2.weight"], w[p + "layer_norm2.bias"], arch.ln_eps)
    h = quick_gelu(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]))
    return x + F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])


def vision_forward(arch: VitArch, w: Dict[str, torch.Tensor], images: torch.Tensor,
                   lora: Optional[Dict[int, Sequence[torch.Tensor]]] = None, scale: float = 2.0,
                   return_hidden: bool = False):
    """images [B,3,S,S] fp32 -> image features [B,P] (un-normalised)."""
    x = vision_embed(arch, w, images)
    hidden = [x] if return_hidden else None
    for i in range(arch.layers):
        x = encoder_layer(arch, w, i, x, lora.get(i) if lora else None, scale)
        if return_hidden:
            hidden.append(x)
    pooled = F.layer_norm(x[:, 0], (arch.width,), w["vision_model.post_layernorm.weight"],
                          w["vision_model.post_layernorm.bias"], arch.ln_eps)
    feats = F.linear(pooled, w["visual_projection.weight"])
    return (feats, hidden) if return_hidden else feats


def clip_logits(feats: torch.Tensor, text_features: torch.Tensor, logit_scale: float) -> torch.Tensor:
    """clip/custom_clip.py:680-687: L2-normalise, scale by exp(logit_scale), dot with unit text feats."""
    f = feats / feats.norm(dim=-1, keepdim=True)
    return math.exp(logit_scale) * f @ text_features.t()


# ----------------------------------------------------------------------------- loss heads
def softmax_entropy(x: torch.Tensor) -> torch.Tensor:
    return -(x.softmax(1) * x.log_softmax(1)).sum(1)


def select_confident_samples(logits: torch.Tensor, top: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """ttl.py:50-54.  The reference's argsort is unstable; the contract here is lowest index first
    among equal entropies."""
    ent = softmax_entropy(logits)
    idx = torch.argsort(ent, descending=False, stable=True)[: int(ent.size(0) * top)]
    return logits[idx], idx


def avg_entropy(outputs: torch.Tensor) -> torch.Tensor:
    logits = outputs - outputs.logsumexp(dim=-1, keepdim=True)
    avg = logits.logsumexp(dim=0) - math.log(logits.shape[0])
    avg = torch.clamp(avg, min=torch.finfo(avg.dtype).min)
    return -(avg * torch.exp(avg)).sum(dim=-1)


def deyo_loss(logits: torch.Tensor, margin_e0: float = 0.4) -> torch.Tensor:
    """deyo.py:97-181 with the script-default flags (filter_ent=0, filter_plpd=0, reweight_ent=1)."""
    ent = softmax_entropy(logits)
    keep = torch.where(ent <= math.log(1000))
    ent = ent[keep]
    coeff = 1.0 / torch.exp(ent.detach() - margin_e0)
    return (ent * coeff).mean(0)


def avg_entropy_grad(sel_logits: torch.Tensor) -> torch.Tensor:
    """Closed form of d avg_entropy / d logits; used to test the CUDA head."""
    lp = sel_logits.log_softmax(-1)
    p = lp.exp()
    a = lp.logsumexp(0) - math.log(sel_logits.shape[0])
    return -(p * (a[None] - (p * a[None]).sum(-1, keepdim=True))) / sel_logits.shape[0]


def deyo_loss_grad(logits: torch.Tensor, margin_e0: float = 0.4) -> torch.Tensor:
    lp = logits.log_softmax(-1)
    p = lp.exp()
    ent = -(p * lp).sum(-1, keepdim=True)
    keep = (ent <= math.log(1000)).to(logits.dtype)
    n = keep.sum().clamp(min=1.0)
    wgt = torch.exp(-(ent - margin_e0)) * keep / n
    return -wgt * p * (lp + ent)


# ----------------------------------------------------------------------------- optimiser
@dataclass
class AdamWState:
    step: int = 0
    m: Dict[Tuple[int, int], torch.Tensor] = field(default_factory=dict)
    v: Dict[Tuple[int, int], torch.Tensor] = field(default_factory=dict)


def adamw_step(params: Dict[int, List[torch.Tensor]], grads: Dict[int, List[Optional[torch.Tensor]]],
               st: AdamWState, lr: float = 5e-3, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
               wd: float = 1e-2) -> None:
    """torch.optim.AdamW (single-tensor path), in place on ``params``."""
    st.step += 1
    bc1 = 1.0 - b1 ** st.step
    bc2 = 1.0 - b2 ** st.step
    for i, plist in params.items():
        for j, p in enumerate(plist):
            g = grads[i][j]
            if g is None:
                continue
            key = (i, j)
            if key not in st.m:
                st.m[key] = torch.zeros_like(p)
                st.v[key] = torch.zeros_like(p)
            p.mul_(1.0 - lr * wd)
            st.m[key].lerp_(g, 1.0 - b1)
            st.v[key].mul_(b2).addcmul_(g, g, value=1.0 - b2)
            denom = (st.v[key].sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(st.m[key], denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------- the loop
@dataclass
class AdaptResult:
    logits0: torch.Tensor                 # [V,C] first-forward logits
    entropies: torch.Tensor               # [V]
    idx: torch.Tensor                     # [K] selected views (head "tpt"); all kept views (head "deyo")
    loss: float
    grads: Dict[int, List[torch.Tensor]]  # of the LAST step
    lora: Dict[int, List[torch.Tensor]]   # post-step factors
    pred_logits: torch.Tensor             # [1,C] adapted prediction on view 0
    losses: List[float] = field(default_factory=list)


def adapt_and_predict(arch: VitArch, w: Dict[str, torch.Tensor], images: torch.Tensor,
                      text_features: torch.Tensor, logit_scale: float, lora0: Dict[int, List[torch.Tensor]],
                      spec: LoraSpec, head: str = "tpt", tta_steps: int = 1, selection_p: float = 0.1,
                      lr: float = 5e-3, margin_e0: float = 0.4,
                      forced_idx: Optional[torch.Tensor] = None) -> AdaptResult:
    """One test sample: reset -> adapt (tta_steps, or tta_steps**2 under the DeYO head)
    -> predict on view 0.  Follows ttl.py:338-352 and ttl.py:70-110 / deyo.py:42-46."""
    lora = {i: [t.clone().requires_grad_(True) for t in ts] for i, ts in lora0.items()}
    st = AdamWState()
    n_opt_steps = tta_steps * tta_steps if head == "deyo" else tta_steps
    sel = forced_idx
    first_logits = None
    ent0 = None
    losses: List[float] = []
    grads: Dict[int, List[torch.Tensor]] = {}
    for _ in range(n_opt_steps):
        feats = vision_forward(arch, w, images, lora, spec.scale)
        logits = clip_logits(feats, text_features, logit_scale)
        if first_logits is None:
            first_logits = logits.detach().clone()
            ent0 = softmax_entropy(first_logits)
        if head == "tpt":
            if sel is None:
                _, sel = select_confident_samples(logits.detach(), selection_p)
            loss = avg_entropy(logits[sel].float())
        else:
            loss = deyo_loss(logits, margin_e0)
            if sel is None:
                sel = torch.where(ent0 <= math.log(1000))[0]
        flat = [t for ts in lora.values() for t in ts]
        gs = torch.autograd.grad(loss, flat, allow_unused=True)
        it = iter(gs)
        grads = {i: [next(it) for _ in ts] for i, ts in lora.items()}
        grads = {i: [torch.zeros_like(p) if g is None else g for g, p in zip(gl, lora[i])] for i, gl in grads.items()}
        with torch.no_grad():
            adamw_step({i: [t for t in ts] for i, ts in lora.items()}, grads, st, lr=lr)
        losses.append(float(loss.detach()))
    with torch.no_grad():
        lora_d = {i: [t.detach() for t in ts] for i, ts in lora.items()}
        feats = vision_forward(arch, w, images[:1], lora_d, spec.scale)
        pred = clip_logits(feats, text_features, logit_scale)
    return AdaptResult(first_logits, ent0, sel, losses[-1] if losses else float("nan"), grads, lora_d, pred, losses)


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)) -> List[torch.Tensor]:
    """utils/tools.py:88-102."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [correct[:k].reshape(-1).float().sum(0, keepdim=True) * (100.0 / target.size(0)) for k in topk]


def make_text_features(n_classes: int, proj: int, seed: int = 11, anchor: Optional[torch.Tensor] = None,
                       margin_boost: float = 0.0) -> torch.Tensor:
    """Unit-norm synthetic text features [C,P].  With ``anchor`` (an image feature) class 0 is pulled
    towards it so the zero-shot top-1 margin is large."""
    g = torch.Generator().manual_seed(seed)
    t = _randn(g, n_classes, proj)
    if anchor is not None and margin_boost > 0:
        a = anchor / anchor.norm()
        t[0] = t[0] + margin_boost * a * t[0].norm()
    return t / t.norm(dim=-1, keepdim=True)
