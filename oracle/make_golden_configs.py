"""Golden fixtures for the BASELINE.json configurations that are benched (configs[1..4]), produced by the UNMODIFIED
reference behind oracle/ref_shim.py where the reference can run the configuration, by the pinned oracle where it cannot.

Dev-container only (needs /root/reference).  Usage:  python oracle/make_golden_configs.py [case ...]

  ref_b16_c1000_tpt : configs[1] -- ViT-B/16, the reference's 1000 ImageNet class names (data/imagnet_prompts.py) through
                      its own prompt builder + text tower, 64 views, north-star head, 1 step.  Reference run.
  ref_b16_c200_tpt  : configs[2] -- the 200 ImageNet-A classes (data/imagenet_variants.py imagenet_a_mask, ttl.py:255-270).
  ref_b16_c200_deyo : the same shape under the script-default head.
  ref_b16_c10_tpt4  : configs[4] -- tta_steps=4 (ttl.py:90-108: selected_idx frozen after step 1), per-step losses recorded
                      by wrapping the reference's avg_entropy (nothing in the reference is edited).
  ref_b16_c10_deyo2 : tta_steps=2 under the DeYO head = 4 optimiser steps, per-step losses via deyo's
                      softmax_entropy wrapper.
  oracle_l14_c10_tpt: configs[3] geometry (ViT-L/14 @224, layers 21-23).  The reference cannot build it (the HF checkpoint
                      name is hard-coded, clip/custom_clip.py:581), so this one comes from oracle/ttl_oracle.py, which the
                      cases above pin to the reference; 64 views, north-star head.

Each fixture stores seeds + small tensors only; weights and views are regenerated from the seeds by the consumer.
"""
from __future__ import annotations

import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402
from oracle import ttl_oracle as O  # noqa: E402

WEIGHT_SEED, LORA_SEED, IMAGE_SEED = 1234, 0, 7
NAMES = ("A_q", "B_q", "A_v", "B_v")
OUT = os.path.join(ROOT, "tests", "golden")


def _classnames(which: str):
    sys.path.insert(0, R.REFERENCE_ROOT)
    from data.imagnet_prompts import imagenet_classes
    if which == "c1000":
        return list(imagenet_classes)
    if which == "c200":
        from data.imagenet_variants import imagenet_a_mask
        return [imagenet_classes[i] for i in imagenet_a_mask]
    return ["airplane", "automobile", "bird", "cat", "deer", "dog", "frog", "horse", "ship", "truck"]


def _save(name, rec):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", flush=True)


def reference_case(name: str, classes: str, head: str, steps: int, image_seed: int = IMAGE_SEED) -> None:
    t0 = time.time()
    arch, spec = O.ARCHS["ViT-B/16"], O.LoraSpec()
    w = O.make_synthetic_weights(arch, WEIGHT_SEED)
    lora0 = O.lora_init(arch, spec, seed=LORA_SEED)
    imgs = O.make_synthetic_views(64, arch.image_size, seed=image_seed)
    ttl_ref, model, opt, optim_state, scaler = R.build_reference_model(w, _classnames(classes))
    R.set_lora(model, lora0)
    with torch.no_grad():
        text = model.get_text_features().clone()
    # the text tower has no trainable part on this path: cache its output so the 1000-class cases finish in minutes
    # (exactly what the reference recomputes at clip/custom_clip.py:667-671; same tensor every time)
    model.get_text_features = lambda: text
    args = R.default_args(deyo_selection=True if head == "deyo" else "", tta_steps=steps)
    with torch.no_grad():
        model.LoRA_reset()
        logits0 = model(imgs).clone()
    opt.load_state_dict(optim_state)
    losses = []
    if head == "tpt":
        orig = ttl_ref.avg_entropy

        def rec_avg_entropy(outputs):
            v = orig(outputs)
            losses.append(float(v.detach()))
            return v
        ttl_ref.avg_entropy = rec_avg_entropy
    # factors the LAST optimiser step starts from (multi-step cases): the gradient of that step can then be checked at the
    # reference's own operating point (B != 0: LoRA branch in the forward, dA, U = dY B) instead of behind three sign-like Adam
    # updates.  Captured by wrapping the optimiser's step at run time; nothing in the reference is edited.
    pre_step = []
    orig_step = opt.step

    def rec_step(*a, **k):
        pre_step.append([[p.detach().clone() for p in R.get_lora(model, spec.layers())[i]] for i in spec.layers()])
        return orig_step(*a, **k)
    opt.step = rec_step
    try:
        ttl_ref.test_time_tuning(model, imgs, opt, scaler, args)
    finally:
        opt.step = orig_step
        if head == "tpt":
            ttl_ref.avg_entropy = orig
    with torch.no_grad():
        pred = model(imgs[:1]).clone()
    _, idx = ttl_ref.select_confident_samples(logits0, args.selection_p)
    ent = -(logits0.softmax(1) * logits0.log_softmax(1)).sum(1)
    lora_now = R.get_lora(model, spec.layers())
    rec = dict(weight_seed=WEIGHT_SEED, lora_seed=LORA_SEED, image_seed=image_seed, logit_scale=np.float64(float(model.logit_scale)),
               tta_steps=steps, head=head, text_features=text.numpy(), logits0=logits0.numpy(), entropies=ent.numpy(),
               idx_sorted=np.sort(idx.numpy()), idx=idx.numpy(), pred_logits=pred.numpy(), losses=np.asarray(losses, dtype=np.float64))
    for i in spec.layers():
        for j, nm in enumerate(NAMES):
            p = lora_now[i][j]
            rec[f"lora_{i}_{nm}"] = p.detach().numpy().copy()
            rec[f"grad_{i}_{nm}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    if head == "tpt" and steps > 1 and len(pre_step) == steps:
        for li, i in enumerate(spec.layers()):
            for j, nm in enumerate(NAMES):
                rec[f"prelast_{i}_{nm}"] = pre_step[-1][li][j].numpy().copy()
    _save(name, rec)
    print(f"  {name}: {time.time() - t0:.0f} s, losses {losses}, optimiser steps seen {len(pre_step)}", flush=True)


def oracle_l14_case(name: str = "oracle_l14_c10_tpt") -> None:
    t0 = time.time()
    arch = O.ARCHS["ViT-L/14"]
    spec = O.LoraSpec(rank=16, alpha=32.0, layer_lo=21, layer_hi=23)
    w = O.make_synthetic_weights(arch, 3)
    lora0 = O.lora_init(arch, spec, 0)
    imgs = O.make_synthetic_views(64, arch.image_size, seed=4)
    text = O.make_text_features(10, arch.proj, seed=5)
    res = O.adapt_and_predict(arch, w, imgs, text, math.log(100.0), lora0, spec, head="tpt")
    rec = dict(weight_seed=3, lora_seed=0, image_seed=4, text_seed=5, logit_scale=np.float64(math.log(100.0)), tta_steps=1,
               head="tpt", text_features=text.numpy(), logits0=res.logits0.numpy(), entropies=res.entropies.numpy(),
               idx_sorted=np.sort(res.idx.numpy()), idx=res.idx.numpy(), pred_logits=res.pred_logits.numpy(),
               losses=np.asarray(res.losses))
    for i in spec.layers():
        for j, nm in enumerate(NAMES):
            rec[f"lora_{i}_{nm}"] = res.lora[i][j].numpy().copy()
            rec[f"grad_{i}_{nm}"] = res.grads[i][j].numpy().copy()
    _save(name, rec)
    print(f"  {name}: {time.time() - t0:.0f} s", flush=True)


CASES = {
    "ref_b16_c1000_tpt": lambda: reference_case("ref_b16_c1000_tpt", "c1000", "tpt", 1),
    "ref_b16_c200_tpt": lambda: reference_case("ref_b16_c200_tpt", "c200", "tpt", 1),
    "ref_b16_c200_deyo": lambda: reference_case("ref_b16_c200_deyo", "c200", "deyo", 1),
    "ref_b16_c10_tpt4": lambda: reference_case("ref_b16_c10_tpt4", "c10", "tpt", 4),
    "ref_b16_c10_deyo2": lambda: reference_case("ref_b16_c10_deyo2", "c10", "deyo", 2),
    "oracle_l14_c10_tpt": oracle_l14_case,
}

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    for c in (sys.argv[1:] or list(CASES)):
        CASES[c]()
