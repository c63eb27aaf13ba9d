"""Host-side handle on one libttl_b200 context (one per process and GPU).

PyTorch is used for device memory and streams only; all arithmetic of the path runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

# geometry per --arch (ttl.py:386); head_dim is 64 for both
ARCH_GEOMETRY = {
    "ViT-B/16": dict(image_size=224, patch=16, width=768, layers=12, heads=12, mlp_dim=3072, proj_dim=512),
    "ViT-L/14": dict(image_size=224, patch=14, width=1024, layers=24, heads=16, mlp_dim=4096, proj_dim=768),
    "ViT-tiny": dict(image_size=64, patch=16, width=128, layers=4, heads=2, mlp_dim=512, proj_dim=64),
}


@dataclass
class Hparams:
    """ttl.py argparse defaults (:367-424) + torch.optim.AdamW defaults (ttl.py:218)."""
    head: str = "tpt"            # "tpt" (deyo_selection falsy) | "deyo" (deyo_selection truthy)
    tta_steps: int = 1
    selection_p: float = 0.1
    lr: float = 5e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 1e-2
    deyo_margin_e0: float = 0.4

    def to_c(self) -> L.TtlHparams:
        return L.TtlHparams(L.HEAD_DEYO if self.head == "deyo" else L.HEAD_TPT, self.tta_steps, self.selection_p,
                            self.lr, self.beta1, self.beta2, self.eps, self.weight_decay, self.deyo_margin_e0)


def _f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


class _DevAlias:
    """Expose library-owned device memory through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 2}


class Pending:
    """Result of an asynchronous host-input call: pinned output tensors that are valid after wait()."""

    def __init__(self, outs, event, keepalive):
        self._outs, self._event, self._keep = outs, event, keepalive

    def wait(self) -> Dict[str, torch.Tensor]:
        self._event.synchronize()
        self._keep = None
        return self._outs


# text towers per --arch (the adapter sits there with --lora_encoder text); `context` tokens per prompt
TEXT_TOWER_GEOMETRY = {
    "ViT-B/16": dict(width=512, layers=12, heads=8, mlp_dim=2048, proj_dim=512, context=77, vocab=49408),
    "ViT-L/14": dict(width=768, layers=12, heads=12, mlp_dim=3072, proj_dim=768, context=77, vocab=49408),
    "tiny": dict(width=128, layers=2, heads=2, mlp_dim=512, proj_dim=64, context=16, vocab=1000),
}


class Engine:
    def __init__(self, arch: str = "ViT-B/16", max_views: int = 64, max_classes: int = 1000, lora_rank: int = 16,
                 lora_alpha: float = 32.0, layer_range: Sequence[int] = (9, 11), device: int = 0,
                 geometry: Optional[dict] = None, max_samples: int = 1, precision: str = "bf16", text_mode: bool = False):
        """text_mode=True: this engine is the CLIP TEXT tower carrying the adapter (`--lora_encoder text`): `max_views` bounds
        the class prompts, `max_classes` the image views per test sample; see set_prompts / adapt_predict_text."""
        if not torch.cuda.is_available():
            raise RuntimeError("ttl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.text_mode = bool(text_mode)
        if self.text_mode:
            g = dict(geometry or TEXT_TOWER_GEOMETRY[arch])
            g.setdefault("image_size", 0)
            g.setdefault("patch", 0)
        else:
            g = dict(geometry or ARCH_GEOMETRY[arch])
        self.arch, self.geom = arch, g
        self.device = torch.device("cuda", device)
        self.max_views, self.max_classes, self.max_samples = max_views, max_classes, max(1, int(max_samples))
        self.rank, self.alpha = lora_rank, lora_alpha
        self.layer_lo, self.layer_hi = int(layer_range[0]), int(layer_range[1])
        self.tokens = g["context"] if self.text_mode else (g["image_size"] // g["patch"]) ** 2 + 1
        self.n_classes = 0
        cfg = L.TtlConfig(g["image_size"], g["patch"], g["width"], g["layers"], g["heads"], g["mlp_dim"], g["proj_dim"],
                          max_views, max_classes, lora_rank, lora_alpha, self.layer_lo, self.layer_hi, 1e-5, device,
                          self.max_samples, {"bf16": L.PRECISION_BF16, "fp32": L.PRECISION_FP32}[precision],
                          1 if self.text_mode else 0, g.get("context", 0), g.get("vocab", 0))
        self.precision = precision
        ctx = C.c_void_p()
        L.check(self.lib.ttl_create(C.byref(ctx), C.byref(cfg)))
        self.ctx = ctx
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.Stream()   # a capturable (non-legacy) stream for the CUDA-graph replay

    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.ttl_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ frozen state
    def _set(self, layer: int, kind: int, arr) -> None:
        a = _f32(arr)
        L.check(self.lib.ttl_set_weight(self.ctx, layer, kind, a.ctypes.data_as(C.c_void_p), a.size), self.ctx)

    def load_weights(self, sd: Dict[str, "torch.Tensor"]) -> None:
        """`sd`: HF CLIP state-dict names of the vision tower + visual_projection (clip/custom_clip.py:581)."""
        p = "vision_model."
        self._set(-1, L.W_CLASS_EMB, sd[p + "embeddings.class_embedding"])
        self._set(-1, L.W_PATCH_EMB, sd[p + "embeddings.patch_embedding.weight"])
        self._set(-1, L.W_POS_EMB, sd[p + "embeddings.position_embedding.weight"])
        self._set(-1, L.W_PRE_LN_G, sd[p + "pre_layrnorm.weight"])
        self._set(-1, L.W_PRE_LN_B, sd[p + "pre_layrnorm.bias"])
        self._set(-1, L.W_POST_LN_G, sd[p + "post_layernorm.weight"])
        self._set(-1, L.W_POST_LN_B, sd[p + "post_layernorm.bias"])
        self._set(-1, L.W_VIS_PROJ, sd["visual_projection.weight"])
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

    # ------------------------------------------------------------------ `--lora_encoder text` (adapter on the text tower)
    def load_text_weights(self, sd: Dict[str, "torch.Tensor"]) -> None:
        """text_mode engine: HF names (`text_model.*`, `text_projection.weight`) as CLIPModel.from_pretrained yields them."""
        p = "text_model."
        self._set(-1, L.W_TOKEN_EMB, sd[p + "embeddings.token_embedding.weight"])
        self._set(-1, L.W_POS_EMB, sd[p + "embeddings.position_embedding.weight"])
        self._set(-1, L.W_POST_LN_G, sd[p + "final_layer_norm.weight"])
        self._set(-1, L.W_POST_LN_B, sd[p + "final_layer_norm.bias"])
        self._set(-1, L.W_VIS_PROJ, sd["text_projection.weight"])
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

    def set_prompts(self, tokens, logit_scale: float) -> None:
        """text_mode engine: tokenised class prompts int [C, context] (clip.tokenize layout); runs the layers below the adapter
        once (reset_classnames, clip/custom_clip.py:343-372)."""
        t = np.ascontiguousarray(tokens.cpu().numpy() if isinstance(tokens, torch.Tensor) else tokens, dtype=np.int32)
        if t.ndim != 2 or t.shape[1] != self.tokens:
            raise ValueError(f"tokens must be [n, {self.tokens}]")
        self.n_classes = int(t.shape[0])
        self._sync_in()
        L.check(self.lib.ttl_text_set_prompts(self.ctx, t.ctypes.data_as(C.c_void_p), t.shape[0], float(logit_scale), self._st()),
                self.ctx)
        self._sync_out()

    def text_features(self) -> torch.Tensor:
        """text_mode engine: L2-normalised class features [C, P] with the current factors (host tensor)."""
        out = np.empty((self.n_classes, self.geom["proj_dim"]), dtype=np.float32)
        L.check(self.lib.ttl_text_features(self.ctx, out.ctypes.data_as(C.c_void_p), self._st()), self.ctx)
        return torch.from_numpy(out)

    def image_features(self, images: torch.Tensor) -> torch.Tensor:
        """Image-tower engine: raw image features [n_views, P] of `images` [n_views,3,S,S] on the device (no adapter, no logits)."""
        images = images.to(self.device, torch.float32).contiguous()
        feats = torch.empty(images.shape[0], self.geom["proj_dim"], device=self.device, dtype=torch.float32)
        self._sync_in()
        L.check(self.lib.ttl_image_features(self.ctx, images.data_ptr(), images.shape[0], feats.data_ptr(), self._st()), self.ctx)
        self._sync_out()
        images.record_stream(self.stream)
        return feats

    def adapt_predict_text(self, img_feats: torch.Tensor, hp: Hparams, forced_idx: Optional[torch.Tensor] = None,
                           want: Sequence[str] = ("pred_logits",)) -> Dict[str, torch.Tensor]:
        """text_mode engine: one test sample with the adapter on the text tower (ttl.py:338-352 with lora_encoder == 'text'):
        `img_feats` [V, P] = image_features() of the sample's views."""
        img_feats = img_feats.to(self.device, torch.float32).contiguous()
        V, n = int(img_feats.shape[0]), self.n_classes
        K = int(V * hp.selection_p)
        outs: Dict[str, torch.Tensor] = {}
        o = L.TtlOutputs()
        shapes = {"logits0": ((V, n), torch.float32), "entropy": ((V,), torch.float32), "idx": ((max(K, 1),), torch.int32),
                  "loss": ((1,), torch.float32), "pred_logits": ((n,), torch.float32)}
        for name in want:
            shp, dt = shapes[name]
            t = torch.empty(shp, dtype=dt, device=self.device)
            outs[name] = t
            setattr(o, name, t.data_ptr())
        fidx = None if forced_idx is None else forced_idx.to(self.device, torch.int32).contiguous()
        h = hp.to_c()
        self._sync_in()
        L.check(self.lib.ttl_text_adapt_predict(self.ctx, img_feats.data_ptr(), V, C.byref(h),
                                                fidx.data_ptr() if fidx is not None else None, C.byref(o), self._st()), self.ctx)
        self._sync_out()
        img_feats.record_stream(self.stream)
        if "idx" in outs:
            outs["idx"] = outs["idx"][:K]
        if "loss" in outs:
            outs["loss"] = outs["loss"][0]
        return outs

    def set_text_features(self, text, logit_scale: float) -> None:
        t = _f32(text)
        self.n_classes = int(t.shape[0])
        L.check(self.lib.ttl_set_text_features(self.ctx, t.ctypes.data_as(C.c_void_p), t.shape[0], t.shape[1],
                                               float(logit_scale)), self.ctx)

    # ------------------------------------------------------------------ adapter
    def set_lora_init(self, lora: Dict[int, Sequence]) -> None:
        for layer, tensors in lora.items():
            for which, t in enumerate(tensors):
                a = _f32(t)
                L.check(self.lib.ttl_lora_set_init(self.ctx, layer, which, a.ctypes.data_as(C.c_void_p), a.size),
                        self.ctx)
        self.lora_reset()

    def lora_reset(self) -> None:
        self._sync_in()
        L.check(self.lib.ttl_lora_reset(self.ctx, self._st()), self.ctx)
        self._sync_out()

    def lora_get(self, layer: int, which: int, what: int = L.LORA_PARAM, sample: int = 0) -> np.ndarray:
        """Factor / gradient / snapshot of one LoRA tensor; `sample` = position in the last concurrent-sample call."""
        d, r = self.geom["width"], self.rank
        shape = (r, d) if which in (L.LORA_A_Q, L.LORA_A_V) else (d, r)
        out = np.empty(shape, dtype=np.float32)
        L.check(self.lib.ttl_lora_get_sample(self.ctx, sample, layer, which, what, out.ctypes.data_as(C.c_void_p), out.size),
                self.ctx)
        return out

    def lora_alias(self, layer: int, which: int, what: int = L.LORA_PARAM) -> torch.Tensor:
        """torch tensor aliasing the library's device buffer (call lora_touch() after writing through it)."""
        ptr, n = C.c_void_p(), C.c_int64()
        L.check(self.lib.ttl_lora_device_ptr(self.ctx, layer, which, what, C.byref(ptr), C.byref(n)), self.ctx)
        d, r = self.geom["width"], self.rank
        shape = (r, d) if which in (L.LORA_A_Q, L.LORA_A_V) else (d, r)
        return torch.as_tensor(_DevAlias(ptr.value, shape), device=self.device)

    def lora_touch(self) -> None:
        # the aliases are written on torch's current stream (optimizer.step / LoRA_AB.reset): order the repack after it
        self._sync_in()
        L.check(self.lib.ttl_lora_touch(self.ctx, self._st()), self.ctx)
        self._sync_out()

    def adamw_step(self, hp: Hparams) -> None:
        h = hp.to_c()
        self._sync_in()
        L.check(self.lib.ttl_adamw_step(self.ctx, C.byref(h), self._st()), self.ctx)
        self._sync_out()

    # ------------------------------------------------------------------ model calls
    def _st(self) -> C.c_void_p:
        return C.c_void_p(self.stream.cuda_stream)

    def _sync_in(self) -> None:
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def _sync_out(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def forward(self, images: torch.Tensor, train: bool = False) -> torch.Tensor:
        """ClipTestTimeTuning.forward: images fp32 [B,3,S,S] on device -> logits fp32 [B,C]."""
        images = images.to(self.device, torch.float32).contiguous()
        logits = torch.empty(images.shape[0], self.n_classes, device=self.device, dtype=torch.float32)
        self._sync_in()
        L.check(self.lib.ttl_forward(self.ctx, images.data_ptr(), images.shape[0], int(train), logits.data_ptr(),
                                     self._st()), self.ctx)
        self._sync_out()
        images.record_stream(self.stream)
        return logits

    def backward(self, dlogits: torch.Tensor) -> None:
        dlogits = dlogits.to(self.device, torch.float32).contiguous()
        self._sync_in()
        L.check(self.lib.ttl_backward(self.ctx, dlogits.data_ptr(), self._st()), self.ctx)
        self._sync_out()
        dlogits.record_stream(self.stream)

    def adapt_predict(self, images: torch.Tensor, hp: Hparams, forced_idx: Optional[torch.Tensor] = None,
                      want: Sequence[str] = ("pred_logits",)) -> Dict[str, torch.Tensor]:
        """One test sample: reset -> adapt -> predict (ttl.py:338-352).  `images` [V,3,S,S] may live on the device (fast
        path) or in (pinned) host memory, in which case the library does the H2D/D2H itself and synchronises."""
        outs = self.adapt_predict_batch(images.unsqueeze(0), hp, None if forced_idx is None else forced_idx.unsqueeze(0), want)
        return {k: v[0] for k, v in outs.items()}

    def adapt_predict_batch(self, images: torch.Tensor, hp: Hparams, forced_idx: Optional[torch.Tensor] = None,
                            want: Sequence[str] = ("pred_logits",), sync: bool = True):
        """S independent test samples adapted concurrently (S <= max_samples), each exactly as adapt_predict would:
        `images` [S,V,3,size,size] -> dict of per-sample results, leading dimension S.
        Host-resident `images` with sync=False return a `Pending` handle instead: the copy of this batch overlaps the
        kernels of the previous one (software-pipelined loader loop); call .wait() for the dict."""
        S, V = int(images.shape[0]), int(images.shape[1])
        if S > self.max_samples:
            raise ValueError(f"{S} samples > max_samples={self.max_samples}")
        K = int(V * hp.selection_p)
        host = not images.is_cuda
        dev = torch.device("cpu") if host else self.device
        images = images.to(torch.float32).contiguous()
        outs: Dict[str, torch.Tensor] = {}
        o = L.TtlOutputs()
        shapes = {"logits0": ((S, V, self.n_classes), torch.float32), "entropy": ((S, V), torch.float32),
                  "idx": ((S, max(K, 1)), torch.int32), "loss": ((S,), torch.float32),
                  "pred_logits": ((S, self.n_classes), torch.float32)}
        for name in want:
            shp, dt = shapes[name]
            t = torch.empty(shp, dtype=dt, device=dev, pin_memory=host)
            outs[name] = t
            setattr(o, name, t.data_ptr())
        fidx = None
        if forced_idx is not None:
            fidx = forced_idx.to(dev, torch.int32).contiguous()
        h = hp.to_c()
        if host:
            fn = self.lib.ttl_adapt_predict_batch_host if sync else self.lib.ttl_adapt_predict_batch_host_async
            L.check(fn(self.ctx, images.data_ptr(), S, V, C.byref(h), fidx.data_ptr() if fidx is not None else None,
                       C.byref(o), self._st()), self.ctx)
            if not sync:
                if "idx" in outs:
                    outs["idx"] = outs["idx"][:, :K]
                ev = torch.cuda.Event()
                ev.record(self.stream)
                return Pending(outs, ev, (images, fidx))
        else:
            self._sync_in()
            L.check(self.lib.ttl_adapt_predict_batch(self.ctx, images.data_ptr(), S, V, C.byref(h),
                                                     fidx.data_ptr() if fidx is not None else None, C.byref(o),
                                                     self._st()), self.ctx)
            self._sync_out()
            images.record_stream(self.stream)
        if "idx" in outs:
            outs["idx"] = outs["idx"][:, :K]
        return outs

    # ------------------------------------------------------------------ optional branches of the weighted-entropy head
    @staticmethod
    def draw_deyo_perms(n_samples: int, n_steps: int, n_kept: int, aug_type: str, patch_len: int, image_size: int):
        """The random draws of deyo.py:127 / :133 from the (CPU) torch RNG, in the order the reference consumes it: one test
        sample after the other, one draw per optimiser step.  patch: argsort(rand(n_kept, patch_len^2)) -> int32
        [S, steps, n_kept, patch_len^2]; pixel: randperm(image_size^2) -> [S, steps, image_size^2]; occ: None."""
        if aug_type == "occ" or n_steps == 0 or n_kept == 0:
            return None
        rows = []
        for _ in range(n_samples):
            for _ in range(n_steps):
                if aug_type == "patch":
                    rows.append(torch.argsort(torch.rand(n_kept, patch_len * patch_len), dim=-1))
                else:
                    rows.append(torch.randperm(image_size * image_size))
        return torch.stack(rows).reshape(n_samples, n_steps, *rows[0].shape).to(torch.int32).contiguous()

    def adapt_predict_batch_deyo(self, images: torch.Tensor, hp: Hparams, filter_ent: int = 0, filter_plpd: int = 0,
                                 reweight_ent: int = 1, reweight_plpd: int = 0, plpd_threshold: float = 0.2,
                                 aug_type: str = "patch", occlusion_size: int = 112, row_start: int = 56, column_start: int = 56,
                                 patch_len: int = 6, perm: Optional[torch.Tensor] = None,
                                 want: Sequence[str] = ("pred_logits",), forced_keep=None) -> Dict[str, torch.Tensor]:
        """adapt_predict_batch with the optional branches of the reference's weighted-entropy head (deyo.py:103-151; flags
        ttl.py:410-424): `images` [S,V,3,size,size] fp32 on the device (x' is built from them).  `perm`: draws for
        aug_type patch / pixel (draw_deyo_perms); drawn here from the torch RNG when omitted.  `forced_keep` [S, n_kept] 0/1
        (parity tests only): teacher-forces the outcome of the PLPD filter."""
        if hp.head != "deyo":
            raise ValueError("the optional DeYO branches belong to head='deyo'")
        S, V = int(images.shape[0]), int(images.shape[1])
        if S > self.max_samples:
            raise ValueError(f"{S} samples > max_samples={self.max_samples}")
        images = images.to(self.device, torch.float32).contiguous()
        n_kept = int(V * hp.selection_p) if filter_ent else V
        n_steps = hp.tta_steps * hp.tta_steps
        aug = {"occ": L.AUG_OCC, "patch": L.AUG_PATCH, "pixel": L.AUG_PIXEL}[aug_type]
        if filter_plpd and perm is None:
            perm = self.draw_deyo_perms(S, n_steps, n_kept, aug_type, patch_len, self.geom["image_size"])
        perm_np = None if (perm is None or not filter_plpd) else np.ascontiguousarray(perm.cpu().numpy(), dtype=np.int32)
        fk = None if forced_keep is None else np.ascontiguousarray(np.asarray(forced_keep), dtype=np.int32).reshape(S, n_kept)
        opt = L.TtlDeyoOptions(int(filter_ent), int(filter_plpd), int(reweight_ent), int(reweight_plpd), float(plpd_threshold), aug,
                               int(occlusion_size), int(row_start), int(column_start), int(patch_len),
                               perm_np.ctypes.data if perm_np is not None else None, perm_np.size if perm_np is not None else 0,
                               fk.ctypes.data if fk is not None else None)
        outs: Dict[str, torch.Tensor] = {}
        o = L.TtlOutputs()
        shapes = {"logits0": ((S, V, self.n_classes), torch.float32), "entropy": ((S, V), torch.float32),
                  "idx": ((S, max(n_kept, 1)), torch.int32), "loss": ((S,), torch.float32),
                  "pred_logits": ((S, self.n_classes), torch.float32)}
        for name in want:
            shp, dt = shapes[name]
            t = torch.empty(shp, dtype=dt, device=self.device)
            outs[name] = t
            setattr(o, name, t.data_ptr())
        h = hp.to_c()
        self._sync_in()
        L.check(self.lib.ttl_adapt_predict_batch_deyo(self.ctx, images.data_ptr(), S, V, C.byref(h), C.byref(opt), C.byref(o),
                                                      self._st()), self.ctx)
        self._sync_out()
        images.record_stream(self.stream)
        if "idx" in outs:
            outs["idx"] = outs["idx"][:, :n_kept] if filter_ent else outs["idx"][:, :0]
        self._last_deyo = (S, n_kept)
        return outs

    def deyo_last_plpd(self):
        """(PLPD values [S, n_kept], final kept-view counts [S]) of the last optimiser step of adapt_predict_batch_deyo."""
        S, n_kept = self._last_deyo
        plpd = np.zeros((S, max(n_kept, 1)), dtype=np.float32)
        n = np.zeros(S, dtype=np.int32)
        L.check(self.lib.ttl_deyo_last_plpd(self.ctx, plpd.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), S, n_kept), self.ctx)
        return plpd[:, :n_kept], n

    # ------------------------------------------------------------------ views generated on the device (views.cu)
    @staticmethod
    def _image_table(images):
        arrs = [np.ascontiguousarray(np.asarray(im, dtype=np.uint8)) for im in images]
        for a in arrs:
            if a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("images must be uint8 [H,W,3]")
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        hs = (C.c_int32 * n)(*[a.shape[0] for a in arrs])
        ws = (C.c_int32 * n)(*[a.shape[1] for a in arrs])
        return arrs, ptrs, hs, ws

    def set_pixel_norm(self, mean, std) -> None:
        m, s = _f32(mean), _f32(std)
        L.check(self.lib.ttl_set_pixel_norm(self.ctx, m.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p)), self.ctx)

    def make_views(self, images, specs) -> torch.Tensor:
        """uint8 images [H_i,W_i,3] + specs [n_images, n_views, 6] (views.ViewSpecSampler) -> fp32 views on the device
        [n_images, n_views, 3, S, S]: what AugMixAugmenter + torch.cat produce on the host (ttl.py:324-336)."""
        from .views import pack_specs
        arrs, ptrs, hs, ws = self._image_table(images)
        sp = pack_specs(specs)
        n, v, size = len(arrs), int(sp.shape[1]), self.geom["image_size"]
        out = torch.empty(n, v, 3, size, size, device=self.device, dtype=torch.float32)
        self._sync_in()
        L.check(self.lib.ttl_make_views(self.ctx, ptrs, hs, ws, n, sp.ctypes.data_as(C.c_void_p), v, out.data_ptr(),
                                        self._st()), self.ctx)
        self._sync_out()
        out.record_stream(self.stream)
        return out

    def adapt_predict_images(self, images, specs, hp: Hparams, forced_idx: Optional[torch.Tensor] = None,
                             want: Sequence[str] = ("pred_logits",), sync: bool = True):
        """adapt_predict_batch fed by decoded uint8 images: the host ships H*W*3 bytes + the view specs per sample, the
        views are generated on the device straight into the patch-embedding operand.  sync=False returns a Pending."""
        from .views import pack_specs
        arrs, ptrs, hs, ws = self._image_table(images)
        sp = pack_specs(specs)
        S, V = len(arrs), int(sp.shape[1])
        if S > self.max_samples:
            raise ValueError(f"{S} samples > max_samples={self.max_samples}")
        if V > self.max_views:
            raise ValueError(f"{V} views > max_views={self.max_views}")
        K = int(V * hp.selection_p)
        outs: Dict[str, torch.Tensor] = {}
        o = L.TtlOutputs()
        shapes = {"logits0": ((S, V, self.n_classes), torch.float32), "entropy": ((S, V), torch.float32),
                  "idx": ((S, max(K, 1)), torch.int32), "loss": ((S,), torch.float32),
                  "pred_logits": ((S, self.n_classes), torch.float32)}
        for name in want:
            shp, dt = shapes[name]
            t = torch.empty(shp, dtype=dt, pin_memory=True)
            outs[name] = t
            setattr(o, name, t.data_ptr())
        fidx = None if forced_idx is None else forced_idx.to("cpu", torch.int32).contiguous()
        h = hp.to_c()
        L.check(self.lib.ttl_adapt_predict_images_async(self.ctx, ptrs, hs, ws, S, sp.ctypes.data_as(C.c_void_p), V,
                                                        C.byref(h), fidx.data_ptr() if fidx is not None else None,
                                                        C.byref(o), self._st()), self.ctx)
        if "idx" in outs:
            outs["idx"] = outs["idx"][:, :K]
        ev = torch.cuda.Event()
        ev.record(self.stream)
        pending = Pending(outs, ev, (arrs, sp, fidx))
        return pending.wait() if sync else pending

    def set_graphs(self, enabled: bool) -> None:
        L.check(self.lib.ttl_set_graphs(self.ctx, int(enabled)), self.ctx)

    def last_launch_count(self) -> int:
        return int(self.lib.ttl_last_launch_count(self.ctx))
