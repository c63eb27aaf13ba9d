"""Generate tests/golden/ref_b16_c10_textlora_{tpt,deyo}.npz by running the UNMODIFIED reference with
`--lora_encoder text` (adapter on the text tower, clip/custom_clip.py:602-606,672-678; ttl.py:146-147,190-191) behind
oracle/ref_shim.py.  Dev-container only (needs /root/reference).  Usage:  python oracle/make_golden_text_lora.py

The fixture pins oracle/text_oracle.py:adapt_and_predict_text_lora -- the checker for the text-tower variant (SURVEY.md 8f
N4), which the CUDA library does not implement yet.  Stored: seeds, the tokenised prompts, first-forward logits, entropies,
selected views, the post-step text-tower factors and their gradients, the adapted prediction (fp32 CPU, ViT-B/16 geometry
on both towers, 10 CIFAR-10 prompts, 64 synthetic views).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402
from oracle import text_oracle as TO  # noqa: E402
from oracle import ttl_oracle as O  # noqa: E402

CIFAR10 = ["airplane", "automobile", "bird", "cat", "deer", "dog", "frog", "horse", "ship", "truck"]
WEIGHT_SEED, TEXT_WEIGHT_SEED, LORA_SEED, IMAGE_SEED = 1234, 4321, 0, 7
LAYERS = range(9, 12)


def run_reference(head: str, steps: int = 1):
    """-> dict of what the reference produced for one sample (also used by tests/test_text_lora_oracle.py)."""
    torch.set_num_threads(os.cpu_count() or 1)
    varch, tarch = O.ARCHS["ViT-B/16"], TO.TEXT_ARCHS["ViT-B/16"]
    weights = {**O.make_synthetic_weights(varch, WEIGHT_SEED), **TO.make_synthetic_text_weights(tarch, TEXT_WEIGHT_SEED)}
    lora0 = TO.text_lora_init(tarch, LAYERS, seed=LORA_SEED)
    imgs = O.make_synthetic_views(64, varch.image_size, seed=IMAGE_SEED)
    ttl_ref, model, opt, optim_state, scaler = R.build_reference_model(weights, CIFAR10, lora_encoder="text")
    R.set_lora(model, lora0, "text")
    args = R.default_args(deyo_selection=True if head == "deyo" else "", tta_steps=steps, lora_encoder="text")
    with torch.no_grad():
        model.LoRA_reset()
        logits0 = model(imgs).clone()
    opt.load_state_dict(optim_state)
    ttl_ref.test_time_tuning(model, imgs, opt, scaler, args)
    with torch.no_grad():
        pred = model(imgs[:1]).clone()
    _, idx = ttl_ref.select_confident_samples(logits0, args.selection_p)
    ent = -(logits0.softmax(1) * logits0.log_softmax(1)).sum(1)
    rec = dict(weight_seed=WEIGHT_SEED, text_weight_seed=TEXT_WEIGHT_SEED, lora_seed=LORA_SEED, image_seed=IMAGE_SEED,
               logit_scale=np.float64(float(model.logit_scale)), tta_steps=steps, head=head,
               tokens=model.prompt_learner.tokenized_prompts.numpy().astype(np.int64),
               logits0=logits0.numpy(), entropies=ent.numpy(), idx_sorted=np.sort(idx.numpy()), pred_logits=pred.numpy())
    now = R.get_lora(model, LAYERS, "text")
    for i in LAYERS:
        for j, nm in enumerate(("A_q", "B_q", "A_v", "B_v")):
            p = now[i][j]
            rec[f"lora_{i}_{nm}"] = p.detach().numpy().copy()
            rec[f"grad_{i}_{nm}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    return rec


def main() -> None:
    outdir = os.path.join(ROOT, "tests", "golden")
    for head in ("tpt", "deyo"):
        rec = run_reference(head)
        path = os.path.join(outdir, f"ref_b16_c10_textlora_{head}.npz")
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
