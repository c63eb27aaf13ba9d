"""Drop-in for the reference's `clip/custom_clip.py` test-time-tuning module API (get_coop, ClipTestTimeTuning, LoRA_AB),
for `--lora_encoder image`, executed by libttl_b200 (hand-written sm_100a CUDA) instead of HF transformers + peft.

What is kept from the reference interface (clip/custom_clip.py:139-217, 570-723):
  * `get_coop(clip_arch, test_set, device, n_ctx, ctx_init, learned_cls, layer_range, init_method, lora_encoder, rank)`
  * `ClipTestTimeTuning(...)` with `.forward(input) -> logits[B,C]`, `.inference`, `.LoRA_reset()`, `.reset()`,
    `.reset_classnames(classnames, arch)`, `.get_text_features()`, `.logit_scale`, `.image_encoder`, `.text_encoder`,
    `.prompt_learner.tokenized_prompts`, `.LoRA_AB`
  * parameter names `image_encoder.vision_model.encoder.layers.{i}.self_attn.{q_proj|v_proj}.lora_{A|B}.default.weight`
    (ttl.py:159-160 string-matches them, ttl.py:197-201 walks them), A [r,d], B [d,r]
  * `LoRA_AB(model, layer_range, init_method, lora_encoder)` with `.init_weights` (one 4-tuple per layer) and `.reset()`

Two execution modes behind that surface:
  * compat: `model(images)` returns logits carrying a grad_fn (one autograd.Function around ttl_forward/ttl_backward), so
    unmodified reference-style code (test_time_tuning, torch.optim.AdamW, GradScaler) can drive it;
  * fast:   `model.adapt_and_predict(images, args)` = one fused C-ABI call / CUDA-graph replay per test sample.

Differences, all documented in DESIGN.md: text features are computed once per class-name set and cached (the reference
re-runs the text tower on every forward); `--rank`/`--arch` are honoured; LoRA in layers outside `layer_range` is exactly
zero in the reference (B=0, never trained) and is exposed as plain tensors that take no part in the computation.
"""
from __future__ import annotations

import math
import os
import zlib
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.init as init

from ttl_b200 import ARCH_GEOMETRY, Engine, Hparams
from ttl_b200 import _lib as L
from ttl_b200.synthetic import synthetic_vit_weights

_HF_NAME = {"ViT-B/16": "openai/clip-vit-base-patch16", "ViT-L/14": "openai/clip-vit-large-patch14"}


# ----------------------------------------------------------------------------- module tree (names only; math is in CUDA)
class _Weight(nn.Module):
    def __init__(self, t: torch.Tensor, trainable: bool):
        super().__init__()
        self.weight = nn.Parameter(t, requires_grad=trainable)


class _LoraProj(nn.Module):
    """Stands where peft's LoRA Linear sits: `.lora_A.default.weight` [r,d], `.lora_B.default.weight` [d,r]."""

    def __init__(self, a: torch.Tensor, b: torch.Tensor, trainable: bool):
        super().__init__()
        self.lora_A = nn.ModuleDict({"default": _Weight(a, trainable)})
        self.lora_B = nn.ModuleDict({"default": _Weight(b, trainable)})


class _SelfAttn(nn.Module):
    def __init__(self, q: _LoraProj, v: _LoraProj):
        super().__init__()
        self.q_proj, self.v_proj = q, v


class _Layer(nn.Module):
    def __init__(self, sa: _SelfAttn):
        super().__init__()
        self.self_attn = sa


class _Encoder(nn.Module):
    def __init__(self, layers: Sequence[_Layer]):
        super().__init__()
        self.layers = nn.ModuleList(layers)


class _VisionModel(nn.Module):
    def __init__(self, enc: _Encoder):
        super().__init__()
        self.encoder = enc


class VisionEncoder(nn.Module):
    """Name-compatible stand-in of the reference's VisionEncoder (clip/custom_clip.py:62-71)."""

    def __init__(self, vm: _VisionModel):
        super().__init__()
        self.vision_model = vm
        self.dtype = torch.float32


class _TextModel(nn.Module):
    def __init__(self, enc: _Encoder):
        super().__init__()
        self.encoder = enc


class PromptEncoder(nn.Module):
    """Text side (clip/custom_clip.py:73-82).  --lora_encoder image: frozen, holds the cached class features.
    --lora_encoder text: carries the adapter module tree `text_model.encoder.layers.{i}.self_attn.{q,v}_proj.lora_{A,B}`
    (ttl.py:146-147,190-191 walk it), aliasing the text-mode engine's factors."""

    def __init__(self, tm: Optional[_TextModel] = None):
        super().__init__()
        if tm is not None:
            self.text_model = tm
        self.dtype = torch.float32


class _PromptState:
    """The part of PromptLearner the image-LoRA path reads (clip/custom_clip.py:655): class names / tokenised prompts."""

    def __init__(self, owner, classnames, ctx_init):
        self.owner = owner
        self.training = False
        self.ctx_init = (ctx_init or "a_photo_of_a").replace("_", " ")
        self.set(classnames)

    def set(self, classnames):
        self.classnames = [c.replace("_", " ") for c in classnames]
        self.prompts = [f"{self.ctx_init} {c}." for c in self.classnames]
        self.tokenized_prompts = None      # filled by _refresh_text_features when a BPE merge table is available

    def reset_classnames(self, classnames, arch):
        self.set(classnames)
        self.owner._refresh_text_features()

    def reset(self):
        pass


# ----------------------------------------------------------------------------- LoRA_AB
class LoRA_AB:
    """clip/custom_clip.py:139-217: initialise A (B stays 0), snapshot, restore the snapshot for layers in range."""

    def __init__(self, model, layer_range, init_method="xavier", lora_encoder="text"):
        self.model = model
        self.layer_range = layer_range
        self.init_method = init_method
        self.lora_encoder = lora_encoder
        self.init_weights = []
        self.initialize_weights()

    def initialize_weights(self):
        if self.init_method in ("xavier", None):
            fn = init.xavier_normal_
        elif self.init_method == "gaussian":
            fn = init.normal_
        elif self.init_method == "kaiming":
            fn = init.kaiming_normal_
        elif self.init_method == "pretrained":
            fn = None
        else:
            raise ValueError(f"Unsupported init_method: {self.init_method}")
        for layer in self._layers():
            self.initialize_layer_weights(layer, fn)

    def initialize_layer_weights(self, layer, fn):
        ws = [layer.self_attn.q_proj.lora_A.default.weight, layer.self_attn.q_proj.lora_B.default.weight,
              layer.self_attn.v_proj.lora_A.default.weight, layer.self_attn.v_proj.lora_B.default.weight]
        if fn is not None:
            with torch.no_grad():
                fn(ws[0])
                fn(ws[2])
        self.init_weights.append(tuple(w.detach().clone() for w in ws))

    def _layers(self):
        """clip/custom_clip.py:163-174: the vision tower's layers, or the text tower's with lora_encoder == 'text'."""
        if self.lora_encoder == "text":
            return self.model.text_model.encoder.layers
        if self.lora_encoder == "image":
            return self.model.vision_model.encoder.layers
        raise NotImplementedError("--lora_encoder prompt (prompt tuning) is outside the TTL path")

    def reset(self):
        layers = self._layers()
        with torch.no_grad():
            for i, layer in enumerate(layers):
                if i in range(self.layer_range[0], self.layer_range[1] + 1):
                    a_q, b_q, a_v, b_v = self.init_weights[i]
                    layer.self_attn.q_proj.lora_A.default.weight.data.copy_(a_q)
                    layer.self_attn.q_proj.lora_B.default.weight.data.copy_(b_q)
                    layer.self_attn.v_proj.lora_A.default.weight.data.copy_(a_v)
                    layer.self_attn.v_proj.lora_B.default.weight.data.copy_(b_v)


# ----------------------------------------------------------------------------- autograd bridge (compat mode)
class _TtlLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, owner, train, *params):
        eng: Engine = owner.engine        # grad mode is off inside Function.forward, so `train` is decided by the caller
        eng.lora_touch()                       # factors may have been written through the aliases (optimizer / reset)
        logits = eng.forward(images, train=train)
        ctx.owner = owner
        ctx.n_params = len(params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        owner = ctx.owner
        owner.engine.backward(dlogits)
        grads = [g.clone() for g in owner._grad_aliases]
        return (None, None, None, *grads)


# ----------------------------------------------------------------------------- the module
class ClipTestTimeTuning(nn.Module):
    def __init__(self, device, classnames, batch_size, criterion="cosine", arch="ViT-B/16", n_ctx=16, ctx_init=None,
                 ctx_position="end", learned_cls=False, layer_range=[9, 11], init_method=None, lora_encoder="text",
                 rank=16, max_views: int = 64, weights: Optional[dict] = None, text_features: Optional[torch.Tensor] = None,
                 logit_scale: Optional[float] = None, max_samples: int = 1, text_weights: Optional[dict] = None,
                 bpe_path: Optional[str] = None, tokenizer=None, precision: str = "bf16", allow_synthetic: bool = False):
        """`weights` / `text_weights` / `logit_scale`: what CLIPModel.from_pretrained hands the reference
        (clip/custom_clip.py:581,619); when `weights` is None the local HF cache is read.  `allow_synthetic` (CLI:
        --synthetic / --random_init, or TTL_SYNTHETIC_WEIGHTS=1) permits seeded random-init towers and prompt-keyed random
        class features when no checkpoint exists; without it a missing checkpoint or text tower raises, as the
        reference's from_pretrained would."""
        super().__init__()
        self.allow_synthetic = bool(allow_synthetic) or os.environ.get("TTL_SYNTHETIC_WEIGHTS", "0") == "1"
        if lora_encoder not in ("image", "text"):
            raise NotImplementedError("--lora_encoder prompt (CoOp/TPT prompt tuning) does not run in the reference either "
                                      "(clip/custom_clip.py:680: image_features unbound); the B200 path implements image and text")
        dev_index = device if isinstance(device, int) else (torch.device(device).index or 0)
        self.device = torch.device("cuda", dev_index)
        self.lora_encoder = lora_encoder
        self.arch = arch
        self.criterion = criterion
        geo = ARCH_GEOMETRY[arch]
        d, n_layers = geo["width"], geo["layers"]
        self.layer_range = [int(layer_range[0]), int(layer_range[1])]
        self.engine = Engine(arch, max_views=max_views, max_classes=max(1000, len(classnames)), lora_rank=rank,
                             lora_alpha=32.0, layer_range=self.layer_range, device=dev_index,
                             max_samples=1 if precision == "fp32" else max_samples, precision=precision)
        self._text_weights, self._bpe_path, self._text_encoder, self._tokenizer = text_weights, bpe_path, None, tokenizer
        hf = None
        if weights is None:
            weights, hf = self._load_vision_weights(arch, self.allow_synthetic)
        self.engine.load_weights(weights)
        if hf is not None:        # the HF checkpoint also carries the text tower and logit_scale (clip/custom_clip.py:619)
            if self._text_weights is None:
                self._text_weights = hf["text"]
            if logit_scale is None:
                logit_scale = hf["logit_scale"]
        if logit_scale is None:
            logit_scale = math.log(100.0)       # CLIP's trained value; random-init runs use it too

        # LoRA module tree.  Trainable-range tensors alias the library's device buffers.
        self._param_aliases, self._grad_aliases = [], []

        def build_layers(eng, n_lyr, width, live):
            out: List[_Layer] = []
            for i in range(n_lyr):
                in_range = live and self.layer_range[0] <= i <= self.layer_range[1]
                ts = []
                for which in range(4):
                    if in_range:
                        t = eng.lora_alias(i, which, L.LORA_PARAM)
                        self._param_aliases.append(t)
                        self._grad_aliases.append(eng.lora_alias(i, which, L.LORA_GRAD))
                    else:
                        shape = (rank, width) if which in (0, 2) else (width, rank)
                        t = torch.zeros(shape, device=self.device)
                    ts.append(t)
                out.append(_Layer(_SelfAttn(_LoraProj(ts[0], ts[1], in_range), _LoraProj(ts[2], ts[3], in_range))))
            return out

        self.text_engine = None
        if lora_encoder == "text":
            # peft wraps the TEXT tower only (clip/custom_clip.py:602-606): a text-mode engine owns the adapter, the image-tower
            # engine above yields the frozen image features of the views
            from ttl_b200.engine import TEXT_TOWER_GEOMETRY
            from ttl_b200.synthetic import HashTokenizer, synthetic_text_weights
            tgeo = TEXT_TOWER_GEOMETRY[arch]
            if self._text_weights is None:
                if not self.allow_synthetic:
                    raise RuntimeError("--lora_encoder text needs the text tower's weights (checkpoint with text_model.* tensors), "
                                       "or --synthetic / --random_init for a seeded random-init tower")
                self._text_weights = synthetic_text_weights(arch, seed=4321)
                if self._tokenizer is None and self._bpe_path is None and not os.environ.get("TTL_BPE_PATH"):
                    self._tokenizer = HashTokenizer(tgeo["vocab"], tgeo["context"])
            self.text_engine = Engine(arch, max_views=max(16, len(classnames)), max_classes=max_views, lora_rank=rank, lora_alpha=32.0,
                                      layer_range=self.layer_range, device=dev_index, text_mode=True)
            self.text_engine.load_text_weights(self._text_weights)
            self.image_encoder = VisionEncoder(_VisionModel(_Encoder(build_layers(self.engine, n_layers, d, False))))
            self.text_encoder = PromptEncoder(_TextModel(_Encoder(build_layers(self.text_engine, tgeo["layers"], tgeo["width"], True))))
            self.LoRA_AB = LoRA_AB(self.text_encoder, layer_range=self.layer_range, init_method=init_method, lora_encoder="text")
        else:
            self.image_encoder = VisionEncoder(_VisionModel(_Encoder(build_layers(self.engine, n_layers, d, True))))
            self.text_encoder = PromptEncoder()
            self.LoRA_AB = LoRA_AB(self.image_encoder, layer_range=self.layer_range, init_method=init_method,
                                   lora_encoder=lora_encoder)
        # snapshot -> library (p0 for the fused reset) ; the live factors already hold it through the aliases
        lora_engine = self.text_engine if lora_encoder == "text" else self.engine
        for i in range(self.layer_range[0], self.layer_range[1] + 1):
            lora_engine.set_lora_init({i: [t.cpu() for t in self.LoRA_AB.init_weights[i]]})
        self.logit_scale = torch.tensor(float(logit_scale), device=self.device)
        self._given_text = text_features
        self.prompt_learner = _PromptState(self, classnames, ctx_init)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self._refresh_text_features()

    # ---- frozen inputs ------------------------------------------------------------------------------
    @staticmethod
    def _load_vision_weights(arch, allow_synthetic: bool):
        """CLIPModel.from_pretrained from the local HF cache (clip/custom_clip.py:581) -> (vision tensors, {text tower,
        logit_scale}).  Without a checkpoint this raises like the reference does, unless random init was asked for."""
        try:
            from transformers import CLIPModel
            m = CLIPModel.from_pretrained(_HF_NAME[arch], local_files_only=True)
        except Exception as e:
            if not allow_synthetic:
                raise RuntimeError(
                    f"no CLIP checkpoint for {arch}: {_HF_NAME.get(arch, arch)} is not in the local HF cache ({type(e).__name__}). "
                    "Pass --vision_checkpoint FILE, or --synthetic N / --random_init for seeded random-init weights") from e
            print("ttl_b200: no local CLIP checkpoint; using seeded random-init ViT weights (random init was requested)")
            return synthetic_vit_weights(arch, seed=1234), None
        sd = m.state_dict()
        vision = {k: v for k, v in sd.items() if k.startswith("vision_model.") or k == "visual_projection.weight"}
        text = {k: v for k, v in sd.items() if k.startswith("text_model.") or k == "text_projection.weight"}
        return vision, {"text": text, "logit_scale": float(m.logit_scale)}

    def _refresh_text_features(self):
        """Once per class-name set (the reference recomputes the text tower in every forward, custom_clip.py:667-671)."""
        names = self.prompt_learner.classnames
        P = ARCH_GEOMETRY[self.arch]["proj_dim"]
        if self.lora_encoder == "text":
            # the class features are a function of the adapter: tokenise, run the layers below the adapter once, keep the rest live
            from ttl_b200.tokenizer import SimpleTokenizer
            if self._tokenizer is None:
                self._tokenizer = SimpleTokenizer(self._bpe_path)
            if len(names) > self.text_engine.max_views:
                raise RuntimeError(f"{len(names)} class prompts > {self.text_engine.max_views} the text engine was sized for")
            self.prompt_learner.tokenized_prompts = self._tokenizer(self.prompt_learner.prompts).to(self.device)
            self.tokenized_prompts = self.prompt_learner.tokenized_prompts
            self.text_engine.set_prompts(self.prompt_learner.tokenized_prompts, float(self.logit_scale))
            self.text_engine.lora_reset()
            self.text_features = self.text_engine.text_features().to(self.device)
            return
        if self._given_text is not None and self._given_text.shape[0] == len(names):
            t = self._given_text.detach().float().cpu()
        elif self._text_weights is not None:
            # the real thing (row N2): tokenise the hand-crafted prompts, run the text tower ONCE on the device
            from ttl_b200.text import TextEncoder
            from ttl_b200.tokenizer import SimpleTokenizer
            if self._tokenizer is None:
                self._tokenizer = SimpleTokenizer(self._bpe_path)
            if self._text_encoder is None:
                self._text_encoder = TextEncoder(self.arch, device=self.device.index or 0)
                self._text_encoder.load_weights(self._text_weights)
            self.prompt_learner.tokenized_prompts = self._tokenizer(self.prompt_learner.prompts).to(self.device)
            self.tokenized_prompts = self.prompt_learner.tokenized_prompts
            t = self._text_encoder.encode(self.prompt_learner.tokenized_prompts)
        else:
            if not self.allow_synthetic:
                raise RuntimeError("no text tower and no class features: the checkpoint has no text_model.* tensors and no "
                                   "text_features= were given (random class features need --synthetic / --random_init)")
            # random-init mode (no text-tower weights given): deterministic unit vectors keyed by the prompt text
            rows = []
            for p in self.prompt_learner.prompts:
                g = torch.Generator().manual_seed(zlib.crc32(p.encode()))
                rows.append(torch.randn(P, generator=g))
            t = torch.stack(rows)
        t = t / t.norm(dim=-1, keepdim=True)
        self.text_features = t.to(self.device)
        self.engine.set_text_features(t, float(self.logit_scale))

    def set_text_features(self, text_features: torch.Tensor, logit_scale: Optional[float] = None):
        """Install externally computed class features [C,P] (e.g. from a real CLIP text tower)."""
        if logit_scale is not None:
            self.logit_scale = torch.tensor(float(logit_scale), device=self.device)
        self._given_text = text_features
        self._refresh_text_features()

    # ---- reference surface -----------------------------------------------------------------------------
    @property
    def dtype(self):
        return torch.float32

    def LoRA_reset(self):
        self.LoRA_AB.reset()

    def reset(self):
        self.prompt_learner.reset()

    def reset_classnames(self, classnames, arch):
        self.prompt_learner.reset_classnames(classnames, arch)

    def get_text_features(self):
        if self.lora_encoder == "text":       # a function of the current adapter state (clip/custom_clip.py:651-663)
            self.text_engine.lora_touch()
            return self.text_engine.text_features().to(self.device)
        return self.text_features

    def _trainable(self):
        return [p for n, p in self.named_parameters() if "lora_" in n and
                any(f"layers.{i}." in n for i in range(self.layer_range[0], self.layer_range[1] + 1))]

    def inference(self, image, label=None, coeff=None):
        if coeff is not None:
            raise NotImplementedError("coeff-weighted feature averaging is not on the TTL path")
        image = image.to(self.device, torch.float32)
        params = self._trainable()
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if self.lora_encoder == "text":
            if train:
                raise NotImplementedError("--lora_encoder text runs on the fused route (adapt_and_predict); --compat needs "
                                          "autograd through the text tower, which the library does not expose")
            f = torch.nn.functional.normalize(self.engine.image_features(image), dim=-1)
            return float(self.logit_scale.exp()) * f @ self.get_text_features().t()
        return _TtlLogits.apply(image, self, train, *params)

    def forward(self, input, label=None, coeff=None):
        if isinstance(input, tuple) or input.dim() == 2:
            raise NotImplementedError("contrastive / directional prompt tuning are not on the TTL path")
        return self.inference(input, label, coeff)

    # ---- fast path -------------------------------------------------------------------------------------
    def hparams_from_args(self, args) -> Hparams:
        deyo = bool(getattr(args, "deyo_selection", True)) and getattr(args, "lora_encoder", "image") != "prompt"
        return Hparams(head="deyo" if deyo else "tpt", tta_steps=int(args.tta_steps), selection_p=float(args.selection_p),
                       lr=float(args.lr), deyo_margin_e0=float(getattr(args, "deyo_margin_e0", 0.4)))

    @staticmethod
    def deyo_general(args) -> bool:
        """True when the weighted-entropy head runs with any of its optional branches (deyo.py:103-151) switched on."""
        deyo = bool(getattr(args, "deyo_selection", True)) and getattr(args, "lora_encoder", "image") != "prompt"
        return deyo and bool(getattr(args, "filter_ent", 0) or getattr(args, "filter_plpd", 0) or getattr(args, "reweight_plpd", 0)
                             or getattr(args, "reweight_ent", 1) != 1)

    def fast_path_ok(self, args) -> bool:
        """Whether one fused library call per batch of samples covers these flags (else: compat mode, the reference's control
        flow over the autograd bridge).  The optional DeYO branches are fused too (ttl_adapt_predict_batch_deyo), on the
        bf16 path; without filter_ent they need C <= 1000 (the ln 1000 filter of deyo.py:107 is then a no-op)."""
        if getattr(args, "cocoop", False):
            return False
        if self.lora_encoder == "text":
            if self.deyo_general(args):
                raise NotImplementedError("--lora_encoder text with the optional DeYO branches (filter_ent / filter_plpd) is not built")
            return True
        if self.deyo_general(args):
            return self.engine.precision == "bf16" and (bool(getattr(args, "filter_ent", 0)) or self.engine.n_classes <= 1000)
        return True

    def adapt_and_predict(self, images: torch.Tensor, args=None, hparams: Optional[Hparams] = None,
                          want=("pred_logits",)):
        """reset -> test_time_tuning -> model(image)  (ttl.py:338-352) as ONE library call.  `images` [V,3,S,S], on the
        device or in pinned host memory.  Returns a dict with `pred_logits` [C] (+ anything else in `want`)."""
        if self.lora_encoder == "text" or (args is not None and hparams is None and self.deyo_general(args)):
            return {k: v[0] for k, v in self.adapt_and_predict_batch(images.unsqueeze(0), args, hparams, want=want).items()}
        hp = hparams or self.hparams_from_args(args)
        return self.engine.adapt_predict(images, hp, want=want)

    def adapt_and_predict_batch(self, images: torch.Tensor, args=None, hparams: Optional[Hparams] = None,
                                want=("pred_logits",), sync: bool = True):
        """The same for S test samples adapted concurrently, each with its own adapter and optimiser state: `images`
        [S,V,3,size,size] (S <= max_samples) -> per-sample results with leading dimension S."""
        hp = hparams or self.hparams_from_args(args)
        if self.lora_encoder == "text":
            # ttl.py:338-352 with the adapter on the text tower: frozen image features of the views, then reset -> adapt ->
            # predict over the class features, one test sample after the other (the class features are per-sample state)
            per = []
            for s_i in range(int(images.shape[0])):
                feats = self.engine.image_features(images[s_i].to(self.device, non_blocking=True))
                per.append(self.text_engine.adapt_predict_text(feats, hp, want=want))
            return {k: torch.stack([p_[k] for p_ in per]) for k in want}
        if args is not None and hparams is None and self.deyo_general(args):
            return self.engine.adapt_predict_batch_deyo(
                images, hp, filter_ent=int(getattr(args, "filter_ent", 0)), filter_plpd=int(getattr(args, "filter_plpd", 0)),
                reweight_ent=int(getattr(args, "reweight_ent", 1)), reweight_plpd=int(getattr(args, "reweight_plpd", 0)),
                plpd_threshold=float(getattr(args, "plpd_threshold", 0.2)), aug_type=str(getattr(args, "aug_type", "patch")),
                occlusion_size=int(getattr(args, "occlusion_size", 112)), row_start=int(getattr(args, "row_start", 56)),
                column_start=int(getattr(args, "column_start", 56)), patch_len=int(getattr(args, "patch_len", 6)), want=want)
        return self.engine.adapt_predict_batch(images, hp, want=want, sync=sync)


    def adapt_and_predict_images(self, images, specs, args=None, hparams: Optional[Hparams] = None,
                                 want=("pred_logits",), sync: bool = True):
        """adapt_and_predict_batch fed by decoded uint8 images [H,W,3] and the view specs the host drew
        (ttl_b200.views.ViewSpecSampler = the RNG half of AugMixAugmenter, data/datautils.py:129-157): the 64 views of
        each sample are generated on the device, bit-exactly as the reference's PIL/torchvision pipeline would."""
        hp = hparams or self.hparams_from_args(args)
        return self.engine.adapt_predict_images(images, specs, hp, want=want, sync=sync)


def get_coop(clip_arch, test_set, device, n_ctx, ctx_init, learned_cls=False, layer_range=[0, 11], init_method=None,
             lora_encoder="text", rank=16, classnames: Optional[Sequence[str]] = None, **kw):
    """clip/custom_clip.py:706-723.  Class names come from the reference's `data` package when importable."""
    if classnames is None:
        classnames = _default_classnames(test_set)
    return ClipTestTimeTuning(device, classnames, None, arch=clip_arch, n_ctx=n_ctx, ctx_init=ctx_init,
                              learned_cls=learned_cls, layer_range=layer_range, init_method=init_method,
                              lora_encoder=lora_encoder, rank=rank, **kw)


def _default_classnames(test_set):
    try:   # the reference's data/ package (class lists are data, out of the hot path; not vendored here)
        from data.imagnet_prompts import imagenet_classes
        from data.fewshot_datasets import fewshot_datasets
        if test_set in fewshot_datasets:
            import data.cls_to_names as c2n
            return getattr(c2n, f"{test_set.lower()}_classes")
        return imagenet_classes
    except Exception:
        return [f"class {i}" for i in range(1000)]
