"""Checkpoint formats -> the library's frozen-weight slots.

The library speaks the HF CLIP vision-tower names that `CLIPModel.from_pretrained` yields (clip/custom_clip.py:581,
`Engine.load_weights`).  The reference also carries OpenAI-format checkpoints (`clip.load`, clip/clip.py:94-140 ->
clip/model.py `VisionTransformer` / `build_model`:428-467): `openai_to_hf_vision` renames/splits those tensors into the HF
layout, `load_vision_checkpoint` reads a .safetensors / .pt / .bin file of either format.  Host-side, once per run."""
from __future__ import annotations

import os
import re
from typing import Dict, NamedTuple, Optional

import torch

_RES = re.compile(r"^visual\.transformer\.resblocks\.(\d+)\.(.+)$")
_BLOCK_MAP = {
    "ln_1.weight": "layer_norm1.weight", "ln_1.bias": "layer_norm1.bias",
    "ln_2.weight": "layer_norm2.weight", "ln_2.bias": "layer_norm2.bias",
    "attn.out_proj.weight": "self_attn.out_proj.weight", "attn.out_proj.bias": "self_attn.out_proj.bias",
    "mlp.c_fc.weight": "mlp.fc1.weight", "mlp.c_fc.bias": "mlp.fc1.bias",
    "mlp.c_proj.weight": "mlp.fc2.weight", "mlp.c_proj.bias": "mlp.fc2.bias",
}
_TOP_MAP = {
    "visual.class_embedding": "vision_model.embeddings.class_embedding",
    "visual.conv1.weight": "vision_model.embeddings.patch_embedding.weight",
    "visual.positional_embedding": "vision_model.embeddings.position_embedding.weight",
    "visual.ln_pre.weight": "vision_model.pre_layrnorm.weight", "visual.ln_pre.bias": "vision_model.pre_layrnorm.bias",
    "visual.ln_post.weight": "vision_model.post_layernorm.weight", "visual.ln_post.bias": "vision_model.post_layernorm.bias",
}


def is_openai_format(sd: Dict[str, torch.Tensor]) -> bool:
    return "visual.conv1.weight" in sd or "visual.proj" in sd


def openai_to_hf_vision(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """OpenAI CLIP state dict (clip/model.py VisionTransformer: fused `in_proj`, `x @ proj`) -> HF names: q/k/v are the
    three row blocks of nn.MultiheadAttention.in_proj_{weight,bias}; visual_projection.weight = proj^T."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        v = v.detach().float()
        if k in _TOP_MAP:
            out[_TOP_MAP[k]] = v
        elif k == "visual.proj":
            out["visual_projection.weight"] = v.t().contiguous()
        else:
            m = _RES.match(k)
            if not m:
                continue          # text tower, logit_scale, ...: not part of the image path
            pre = f"vision_model.encoder.layers.{m.group(1)}."
            name = m.group(2)
            if name in ("attn.in_proj_weight", "attn.in_proj_bias"):
                d = v.shape[0] // 3
                suffix = "weight" if name.endswith("weight") else "bias"
                for j, p in enumerate(("q_proj", "k_proj", "v_proj")):
                    out[f"{pre}self_attn.{p}.{suffix}"] = v[j * d:(j + 1) * d].contiguous()
            elif name in _BLOCK_MAP:
                out[pre + _BLOCK_MAP[name]] = v
    if "vision_model.embeddings.patch_embedding.weight" not in out:
        raise KeyError("not an OpenAI-format CLIP state dict (no visual.conv1.weight)")
    return out


def hf_vision_subset(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: v.detach().float() for k, v in sd.items() if k.startswith("vision_model.") or k == "visual_projection.weight"}


def _read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(path)
    else:
        try:
            sd = torch.jit.load(path, map_location="cpu").state_dict()      # OpenAI releases are TorchScript archives
        except Exception:
            sd = torch.load(path, map_location="cpu", weights_only=True)
            if isinstance(sd, dict) and "state_dict" in sd:
                sd = sd["state_dict"]
    return sd


def load_vision_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """Read a CLIP checkpoint file (HF `model.safetensors` / `pytorch_model.bin`, or an OpenAI `ViT-B-16.pt` -- a TorchScript
    archive or a plain state dict) and return the HF-named vision tensors `Engine.load_weights` takes."""
    sd = _read_state_dict(path)
    return openai_to_hf_vision(sd) if is_openai_format(sd) else hf_vision_subset(sd)


class ClipCheckpoint(NamedTuple):
    vision: Dict[str, torch.Tensor]            # HF names: vision_model.*, visual_projection.weight
    text: Optional[Dict[str, torch.Tensor]]    # HF names: text_model.*, text_projection.weight; None if the file has no text tower
    logit_scale: Optional[float]               # log domain, as CLIPModel.logit_scale (clip/custom_clip.py:619)
    bpe_path: Optional[str]                    # merge table found next to the file (HF `merges.txt` or the OpenAI .txt.gz)


def load_clip_checkpoint(path: str) -> ClipCheckpoint:
    """Everything the TTL path needs from one CLIP checkpoint file: the image tower, the text tower that turns the class
    prompts into the cached class features (clip/custom_clip.py:651-663) and `logit_scale` (:619) -- what
    `CLIPModel.from_pretrained` (clip/custom_clip.py:581) hands the reference in one object."""
    from .text import openai_to_hf_text
    sd = _read_state_dict(path)
    if is_openai_format(sd):
        vision = openai_to_hf_vision(sd)
        text = openai_to_hf_text(sd) if "token_embedding.weight" in sd else None
    else:
        vision = hf_vision_subset(sd)
        text = {k: v.detach().float() for k, v in sd.items() if k.startswith("text_model.") or k == "text_projection.weight"}
        text = text if "text_projection.weight" in text else None
    scale = float(sd["logit_scale"]) if "logit_scale" in sd else None
    here = os.path.dirname(os.path.abspath(path))
    bpe = next((os.path.join(here, n) for n in ("merges.txt", "bpe_simple_vocab_16e6.txt.gz")
                if os.path.exists(os.path.join(here, n))), None)
    return ClipCheckpoint(vision, text, scale, bpe)
