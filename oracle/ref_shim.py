"""Run the UNMODIFIED reference (/root/reference) on CPU behind import shims.

TEST INFRASTRUCTURE ONLY, and dev-container only: /root/reference does not exist on the GPU box,
so nothing that runs there imports this file.  It is used by ``oracle/make_golden.py`` to produce
``tests/golden/*.npz`` and by ``tests/test_oracle_vs_reference.py`` (skipped when the reference is
absent) to pin ``oracle/ttl_oracle.py`` against the reference's own code.

What is shimmed (nothing in the reference's files is edited or copied):
  * ``ftfy`` (clip/simple_tokenizer.py:6), ``matplotlib`` (deyo.py:14): absent here -> stub modules.
  * ``peft`` (clip/custom_clip.py:567): absent here -> minimal stand-in with peft's documented LoRA
    ``Linear`` semantics  y = base(x) + lora_B(lora_A(dropout(x))) * alpha/r ; A kaiming-uniform,
    B zeros; all non-LoRA parameters frozen.
  * ``CLIPModel.from_pretrained`` (clip/custom_clip.py:581): no network/weights -> random-init
    ``CLIPModel`` whose vision tower is then overwritten with the oracle's seeded synthetic weights;
    ``get_image_features/get_text_features`` wrapped to return the pooled tensor (transformers 5.x
    returns an output object; the reference was written against 4.x which returned the tensor).
  * ``clip.load`` (clip/custom_clip.py:580): OpenAI checkpoint download -> random-init
    ``clip.model.CLIP`` (only its token embedding/tokeniser is used on this path).
"""
from __future__ import annotations

import importlib
import importlib.util
import math
import os
import sys
import types
from typing import Dict, List, Optional

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("TTL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ttl.py"))


# ----------------------------------------------------------------------------- peft stand-in
class _LoraLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, alpha: float, dropout: float):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        self.lora_dropout = nn.ModuleDict({"default": nn.Dropout(dropout) if dropout > 0 else nn.Identity()})
        self.scaling = alpha / r
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x):
        y = self.base_layer(x)
        d = self.lora_dropout["default"](x)
        return y + self.lora_B["default"](self.lora_A["default"](d)) * self.scaling


def _make_peft_module() -> types.ModuleType:
    peft = types.ModuleType("peft")

    class LoraConfig:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class TaskType:
        FEATURE_EXTRACTION = "FEATURE_EXTRACTION"

    def prepare_model_for_int8_training(model, *a, **k):
        for p in model.parameters():
            p.requires_grad_(False)
            if p.dtype in (torch.float16, torch.bfloat16):
                p.data = p.data.float()
        return model

    class _PeftWrapper(nn.Module):
        def __init__(self, model):
            super().__init__()
            self.base_model = types.SimpleNamespace(model=model)

    def get_peft_model(model, cfg):
        targets = set(cfg.target_modules)
        for parent in list(model.modules()):
            for name, child in list(parent.named_children()):
                if name in targets and isinstance(child, nn.Linear):
                    setattr(parent, name, _LoraLinear(child, cfg.r, cfg.lora_alpha, cfg.lora_dropout))
        for n, p in model.named_parameters():
            p.requires_grad_("lora_" in n)
        return _PeftWrapper(model)

    peft.LoraConfig = LoraConfig
    peft.TaskType = TaskType
    peft.prepare_model_for_int8_training = prepare_model_for_int8_training
    peft.get_peft_model = get_peft_model
    return peft


# ----------------------------------------------------------------------------- installation
_STATE: Dict[str, object] = {}


def install(vision_weights: Optional[Dict[str, torch.Tensor]] = None, logit_scale: float = math.log(100.0),
            seed: int = 1234):
    """Install the shims and import the reference.  Returns ``(ttl_module, clip_package)``."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    _STATE["vision_weights"] = vision_weights
    _STATE["logit_scale"] = logit_scale
    _STATE["seed"] = seed
    if "ttl_ref" in _STATE:
        return _STATE["ttl_ref"], _STATE["clip_pkg"]

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for mod in ("clip", "data", "utils", "deyo"):
        if mod in sys.modules and not getattr(sys.modules[mod], "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[mod]

    ftfy = types.ModuleType("ftfy")
    ftfy.fix_text = lambda s: s
    sys.modules.setdefault("ftfy", ftfy)
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    sys.modules["peft"] = _make_peft_module()

    import transformers
    from transformers import CLIPConfig, CLIPModel

    def _from_pretrained(name, *a, **k):
        torch.manual_seed(int(_STATE["seed"]))
        cfg = CLIPConfig(vision_config={"patch_size": 16}, logit_scale_init_value=float(_STATE["logit_scale"]))
        try:
            cfg._attn_implementation = "eager"
        except Exception:
            pass
        model = CLIPModel(cfg).eval()
        vw = _STATE.get("vision_weights")
        if vw is not None:
            sd = model.state_dict()
            for k_, v_ in vw.items():
                assert k_ in sd and sd[k_].shape == v_.shape, k_
            model.load_state_dict({**sd, **{k_: v_.clone() for k_, v_ in vw.items()}})
        gif, gtf = model.get_image_features, model.get_text_features

        def _pooled(fn):
            def wrapped(*aa, **kk):
                out = fn(*aa, **kk)
                return out if torch.is_tensor(out) else out.pooler_output
            return wrapped

        model.get_image_features = _pooled(gif)
        model.get_text_features = _pooled(gtf)
        return model

    CLIPModel.from_pretrained = staticmethod(_from_pretrained)

    clip_pkg = importlib.import_module("clip")
    clip_model_mod = importlib.import_module("clip.model")

    def _load(name, device="cpu", download_root=None, **k):
        torch.manual_seed(int(_STATE["seed"]) + 1)
        m = clip_model_mod.CLIP(512, 224, 12, 768, 16, 77, 49408, 512, 8, 12).eval().float()
        return m, 512, None

    for modname in ("clip", "clip.clip", "clip.custom_clip"):
        m = sys.modules.get(modname)
        if m is not None and hasattr(m, "load"):
            setattr(m, "load", _load)

    spec = importlib.util.spec_from_file_location("ttl_ref", os.path.join(REFERENCE_ROOT, "ttl.py"))
    ttl_ref = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = ["ttl.py"]
    try:
        spec.loader.exec_module(ttl_ref)
    finally:
        sys.argv = argv
    sys.modules["ttl_ref"] = ttl_ref
    _STATE["ttl_ref"] = ttl_ref
    _STATE["clip_pkg"] = clip_pkg
    return ttl_ref, clip_pkg


def default_args(**over) -> types.SimpleNamespace:
    """The argparse defaults of ttl.py:367-424 that the path reads."""
    a = dict(cocoop=False, deyo_selection=True, lora_encoder="image", tta_steps=1, selection_p=0.1, lr=5e-3,
             deyo_margin=0.5, deyo_margin_e0=0.4, filter_ent=0, filter_plpd=0, reweight_ent=1, reweight_plpd=0,
             aug_type="patch", occlusion_size=112, patch_len=6, row_start=56, column_start=56,
             plpd_threshold=0.2, layer_range=[9, 11], gpu=0, arch="ViT-B/16", init_method="xavier", rank=16,
             n_ctx=4, ctx_init="a_photo_of_a", test_sets="A", print_freq=10, tpt=True)
    a.update(over)
    return types.SimpleNamespace(**a)


def _lora_layers(model, lora_encoder: str):
    """The layer list ttl.py:190-193 walks for the optimizer groups."""
    if lora_encoder == "text":
        return model.text_encoder.text_model.encoder.layers
    return model.image_encoder.vision_model.encoder.layers


def build_reference_model(vision_weights, classnames: List[str], lora_seed: int = 0, layer_range=(9, 11),
                          lora_encoder: str = "image"):
    """get_coop(...) -> requires_grad filter (ttl.py:151-163) -> AdamW param groups (ttl.py:189-220).
    `vision_weights` may also carry `text_model.*` / `text_projection.weight` entries (same CLIPModel state dict);
    `lora_encoder='text'` builds the reference's text-tower variant (clip/custom_clip.py:602-606, ttl.py:146-147,190-191)."""
    from copy import deepcopy
    ttl_ref, clip_pkg = install(vision_weights)
    from clip.custom_clip import get_coop
    torch.manual_seed(lora_seed)
    model = get_coop("ViT-B/16", "A", "cpu", 4, "a_photo_of_a", layer_range=list(layer_range),
                     init_method="xavier", lora_encoder=lora_encoder, rank=16)
    model.reset_classnames(classnames, "ViT-B/16")
    enc_name = "text_encoder" if lora_encoder == "text" else "image_encoder"
    for name, p in model.named_parameters():
        ok = (enc_name in name and ("lora_A" in name or "lora_B" in name)
              and any(f"layers.{i}." in name for i in range(layer_range[0], layer_range[1] + 1)))
        p.requires_grad_(ok)
    groups = []
    for i, layer in enumerate(_lora_layers(model, lora_encoder)):
        if layer_range[0] <= i <= layer_range[1]:
            groups.extend([{"params": layer.self_attn.q_proj.lora_A.parameters()},
                           {"params": layer.self_attn.q_proj.lora_B.parameters()},
                           {"params": layer.self_attn.v_proj.lora_A.parameters()},
                           {"params": layer.self_attn.v_proj.lora_B.parameters()}])
    opt = torch.optim.AdamW(groups, lr=5e-3)
    optim_state = deepcopy(opt.state_dict())
    scaler = torch.amp.GradScaler("cuda", init_scale=1000, enabled=False)
    model.eval()
    return ttl_ref, model, opt, optim_state, scaler


def set_lora(model, lora: Dict[int, List[torch.Tensor]], lora_encoder: str = "image") -> None:
    """Overwrite the reference model's LoRA factors AND its reset snapshot with given tensors."""
    layers = _lora_layers(model, lora_encoder)
    with torch.no_grad():
        for i, (a_q, b_q, a_v, b_v) in lora.items():
            sa = layers[i].self_attn
            sa.q_proj.lora_A.default.weight.copy_(a_q)
            sa.q_proj.lora_B.default.weight.copy_(b_q)
            sa.v_proj.lora_A.default.weight.copy_(a_v)
            sa.v_proj.lora_B.default.weight.copy_(b_v)
            model.LoRA_AB.init_weights[i] = (a_q.clone(), b_q.clone(), a_v.clone(), b_v.clone())


def get_lora(model, layers_range, lora_encoder: str = "image") -> Dict[int, List[torch.Tensor]]:
    layers = _lora_layers(model, lora_encoder)
    out = {}
    for i in layers_range:
        sa = layers[i].self_attn
        out[i] = [sa.q_proj.lora_A.default.weight, sa.q_proj.lora_B.default.weight,
                  sa.v_proj.lora_A.default.weight, sa.v_proj.lora_B.default.weight]
    return out
