"""Generate tests/golden/*.npz by running the UNMODIFIED reference (behind oracle/ref_shim.py).

Dev-container only (needs /root/reference).  Usage:  python oracle/make_golden.py
Each fixture stores only seeds + small tensors; the 86 M synthetic weights and the 64 views are
regenerated from the seeds by ``oracle.ttl_oracle`` wherever the fixture is consumed.

Cases (BASELINE.json configs[0]: CLIP ViT-B/16 random-init, 10 CIFAR-10 class prompts, 64 synthetic
224x224 views, LoRA r=16, fp32 CPU):
  ref_b16_c10_tpt   : north-star head (ttl.py:86-110), 1 step
  ref_b16_c10_deyo  : script-default weighted-entropy head (deyo.py:93-196), 1 step
  ref_b16_c10_tpt2  : north-star head, 2 steps (dA != 0, selected_idx reuse at ttl.py:97-98)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402
from oracle import ttl_oracle as O  # noqa: E402

CIFAR10 = ["airplane", "automobile", "bird", "cat", "deer", "dog", "frog", "horse", "ship", "truck"]
WEIGHT_SEED, LORA_SEED, IMAGE_SEED = 1234, 0, 7


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    arch = O.ARCHS["ViT-B/16"]
    spec = O.LoraSpec()
    w = O.make_synthetic_weights(arch, WEIGHT_SEED)
    lora0 = O.lora_init(arch, spec, seed=LORA_SEED)
    imgs = O.make_synthetic_views(64, arch.image_size, seed=IMAGE_SEED)
    ttl_ref, model, opt, optim_state, scaler = R.build_reference_model(w, CIFAR10)
    R.set_lora(model, lora0)
    with torch.no_grad():
        text = model.get_text_features().clone()
    logit_scale = float(model.logit_scale)
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)

    for case, head, steps in (("tpt", "tpt", 1), ("deyo", "deyo", 1), ("tpt2", "tpt", 2)):
        args = R.default_args(deyo_selection=True if head == "deyo" else "", tta_steps=steps)
        with torch.no_grad():
            model.LoRA_reset()
            logits0 = model(imgs).clone()
        opt.load_state_dict(optim_state)
        ttl_ref.test_time_tuning(model, imgs, opt, scaler, args)
        with torch.no_grad():
            pred = model(imgs[:1]).clone()
        _, idx = ttl_ref.select_confident_samples(logits0, args.selection_p)
        ent = -(logits0.softmax(1) * logits0.log_softmax(1)).sum(1)
        lora_now = R.get_lora(model, spec.layers())
        rec = dict(weight_seed=WEIGHT_SEED, lora_seed=LORA_SEED, image_seed=IMAGE_SEED,
                   logit_scale=np.float64(logit_scale), tta_steps=steps, head=head,
                   text_features=text.numpy(), logits0=logits0.numpy(), entropies=ent.numpy(),
                   idx_sorted=np.sort(idx.numpy()), pred_logits=pred.numpy())
        for i in spec.layers():
            for j, nm in enumerate(("A_q", "B_q", "A_v", "B_v")):
                p = lora_now[i][j]
                rec[f"lora_{i}_{nm}"] = p.detach().numpy().copy()
                rec[f"grad_{i}_{nm}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        path = os.path.join(outdir, f"ref_b16_c10_{case}.npz")
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
