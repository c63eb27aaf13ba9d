"""Class-feature builder on the device: host handle on a libttl_b200 text-tower context.

`TextEncoder.encode(tokens)` is what `ClipTestTimeTuning.get_text_features()` (clip/custom_clip.py:651-663) returns --
L2-normalised class features [C, P] -- computed once per class-name set instead of inside every forward."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _lib as L

TEXT_GEOMETRY = {
    "ViT-B/16": dict(vocab=49408, context=77, width=512, layers=12, heads=8, mlp_dim=2048, proj_dim=512),
    "ViT-L/14": dict(vocab=49408, context=77, width=768, layers=12, heads=12, mlp_dim=3072, proj_dim=768),
    "tiny": dict(vocab=1000, context=16, width=128, layers=2, heads=2, mlp_dim=512, proj_dim=64),
}


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def openai_to_hf_text(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """OpenAI CLIP text-tower names (clip/model.py CLIP: token_embedding, positional_embedding, transformer.resblocks.*,
    ln_final, text_projection [d,P] applied as x @ W) -> HF names (text_projection.weight = W^T)."""
    out: Dict[str, torch.Tensor] = {}
    ren = {"ln_1": "layer_norm1", "ln_2": "layer_norm2", "attn.out_proj": "self_attn.out_proj", "mlp.c_fc": "mlp.fc1",
           "mlp.c_proj": "mlp.fc2"}
    for k, v in sd.items():
        v = v.detach().float()
        if k == "token_embedding.weight":
            out["text_model.embeddings.token_embedding.weight"] = v
        elif k == "positional_embedding":
            out["text_model.embeddings.position_embedding.weight"] = v
        elif k in ("ln_final.weight", "ln_final.bias"):
            out["text_model.final_layer_norm." + k.split(".")[1]] = v
        elif k == "text_projection":
            out["text_projection.weight"] = v.t().contiguous()
        elif k.startswith("transformer.resblocks."):
            _, _, i, rest = k.split(".", 3)
            pre = f"text_model.encoder.layers.{i}."
            mod, suffix = rest.rsplit(".", 1) if rest.count(".") else (rest, "")
            if rest.startswith("attn.in_proj_"):
                d = v.shape[0] // 3
                suffix = "weight" if rest.endswith("weight") else "bias"
                for j, p in enumerate(("q_proj", "k_proj", "v_proj")):
                    out[f"{pre}self_attn.{p}.{suffix}"] = v[j * d:(j + 1) * d].contiguous()
            elif mod in ren:
                out[f"{pre}{ren[mod]}.{suffix}"] = v
    if "text_model.embeddings.token_embedding.weight" not in out:
        raise KeyError("not an OpenAI-format CLIP state dict (no token_embedding.weight)")
    return out


class TextEncoder:
    def __init__(self, arch: str = "ViT-B/16", device: int = 0, max_prompts: int = 256, geometry: dict | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("ttl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        g = dict(geometry or TEXT_GEOMETRY[arch])
        self.geom = g
        cfg = L.TtlTextConfig(g["vocab"], g["context"], g["width"], g["layers"], g["heads"], g["mlp_dim"], g["proj_dim"],
                              max_prompts, 1e-5, device)
        ctx = C.c_void_p()
        rc = self.lib.ttl_text_create(C.byref(ctx), C.byref(cfg))
        if rc != 0:
            raise RuntimeError(f"libttl_b200 error {rc}: {(self.lib.ttl_text_last_error(None) or b'').decode()}")
        self.ctx = ctx
        self.device = torch.device("cuda", device)

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"libttl_b200 error {rc}: {(self.lib.ttl_text_last_error(self.ctx) or b'').decode()}")

    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.ttl_text_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _set(self, layer: int, kind: int, arr) -> None:
        a = _f32(arr)
        self._check(self.lib.ttl_text_set_weight(self.ctx, layer, kind, a.ctypes.data_as(C.c_void_p), a.size))

    def load_weights(self, sd: Dict[str, torch.Tensor]) -> None:
        """HF names (`text_model.*`, `text_projection.weight`) as CLIPModel.from_pretrained yields them."""
        p = "text_model."
        self._set(-1, L.TW_TOKEN_EMB, sd[p + "embeddings.token_embedding.weight"])
        self._set(-1, L.TW_POS_EMB, sd[p + "embeddings.position_embedding.weight"])
        self._set(-1, L.TW_FINAL_LN_G, sd[p + "final_layer_norm.weight"])
        self._set(-1, L.TW_FINAL_LN_B, sd[p + "final_layer_norm.bias"])
        self._set(-1, L.TW_TEXT_PROJ, sd["text_projection.weight"])
        for i in range(self.geom["layers"]):
            q = f"{p}encoder.layers.{i}."
            for kind, name in ((L.W_LN1_G, "layer_norm1.weight"), (L.W_LN1_B, "layer_norm1.bias"),
                               (L.W_Q_W, "self_attn.q_proj.weight"), (L.W_Q_B, "self_attn.q_proj.bias"),
                               (L.W_K_W, "self_attn.k_proj.weight"), (L.W_K_B, "self_attn.k_proj.bias"),
                               (L.W_V_W, "self_attn.v_proj.weight"), (L.W_V_B, "self_attn.v_proj.bias"),
                               (L.W_O_W, "self_attn.out_proj.weight"), (L.W_O_B, "self_attn.out_proj.bias"),
                               (L.W_LN2_G, "layer_norm2.weight"), (L.W_LN2_B, "layer_norm2.bias"),
                               (L.W_FC1_W, "mlp.fc1.weight"), (L.W_FC1_B, "mlp.fc1.bias"),
                               (L.W_FC2_W, "mlp.fc2.weight"), (L.W_FC2_B, "mlp.fc2.bias")):
                self._set(i, kind, sd[q + name])

    def encode(self, tokens) -> torch.Tensor:
        """tokens int [C, context] (clip.tokenize layout) -> L2-normalised class features fp32 [C, P] (host tensor)."""
        t = np.ascontiguousarray(tokens.cpu().numpy() if isinstance(tokens, torch.Tensor) else tokens, dtype=np.int32)
        if t.ndim != 2 or t.shape[1] != self.geom["context"]:
            raise ValueError(f"tokens must be [n, {self.geom['context']}]")
        out = np.empty((t.shape[0], self.geom["proj_dim"]), dtype=np.float32)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            self._check(self.lib.ttl_text_encode(self.ctx, t.ctypes.data_as(C.c_void_p), t.shape[0],
                                                 out.ctypes.data_as(C.c_void_p), C.c_void_p(st)))
        return torch.from_numpy(out)
