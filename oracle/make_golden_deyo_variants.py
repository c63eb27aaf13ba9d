"""Golden fixtures for the non-default branches of the reference's weighted-entropy head:
deyo.py:103-151 `filter_ent` (top-p selection inside the DeYO head) and `filter_plpd` (second forward on a
structure-destroyed copy of the kept views, keep those whose top-class probability drops by more than plpd_threshold).

Dev-container only: runs the UNMODIFIED reference (behind oracle/ref_shim.py).  Usage: python oracle/make_golden_deyo_variants.py
Fixtures (ViT-B/16 random-init, 10 CIFAR-10 prompts, 64 synthetic views, fp32 CPU; weights/views regenerate from seeds):
  ref_b16_c10_deyo_fent        filter_ent=1
  ref_b16_c10_deyo_plpd_occ    filter_plpd=1, aug_type='occ'    (deterministic occlusion window)
  ref_b16_c10_deyo_plpd_patch  filter_plpd=1, aug_type='patch'  (6x6 patch shuffle; permutation from the CPU torch RNG, seed stored)
plpd_threshold is placed in the widest gap of the reference's own PLPD values near their median, so that bf16 noise on the
device cannot flip a view across it; the chosen value is stored and passed back through args in the test."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402
from oracle import ttl_oracle as O  # noqa: E402
from oracle.make_golden import CIFAR10, WEIGHT_SEED, LORA_SEED, IMAGE_SEED  # noqa: E402

RNG_SEED = 123


def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    arch, spec = O.ARCHS["ViT-B/16"], O.LoraSpec()
    w = O.make_synthetic_weights(arch, WEIGHT_SEED)
    lora0 = O.lora_init(arch, spec, seed=LORA_SEED)
    imgs = O.make_synthetic_views(64, arch.image_size, seed=IMAGE_SEED)
    ttl_ref, model, opt, optim_state, scaler = R.build_reference_model(w, CIFAR10)
    R.set_lora(model, lora0)
    import deyo as deyo_ref   # the reference's module (ref_shim put /root/reference on sys.path)
    with torch.no_grad():
        text = model.get_text_features().clone()
    outdir = os.path.join(ROOT, "tests", "golden")

    for case, over in (("fent", dict(filter_ent=1)),
                       ("plpd_occ", dict(filter_plpd=1, aug_type="occ")),
                       ("plpd_patch", dict(filter_plpd=1, aug_type="patch"))):
        args = R.default_args(deyo_selection=True, tta_steps=1, **over)
        thr = None
        if over.get("filter_plpd"):
            # dry run with a threshold that keeps everything, recording p(top) - p'(top) through a hook on softmax outputs
            vals = _reference_plpd(deyo_ref, model, imgs, args, opt, optim_state, scaler)
            s = np.sort(vals)
            lo, hi = len(s) // 4, 3 * len(s) // 4
            gaps = s[lo + 1:hi + 1] - s[lo:hi]
            k = int(np.argmax(gaps)) + lo
            thr = float(0.5 * (s[k] + s[k + 1]))
            args.plpd_threshold = thr
            print(case, "plpd range", s[0], s[-1], "threshold", thr, "gap", float(gaps.max()), "kept", int((vals > thr).sum()))
        with torch.no_grad():
            model.LoRA_reset()
            logits0 = model(imgs).clone()
        opt.load_state_dict(optim_state)
        torch.manual_seed(RNG_SEED)
        ttl_ref.test_time_tuning(model, imgs, opt, scaler, args)
        with torch.no_grad():
            pred = model(imgs[:1]).clone()
        lora_now = R.get_lora(model, spec.layers())
        rec = dict(weight_seed=WEIGHT_SEED, lora_seed=LORA_SEED, image_seed=IMAGE_SEED, rng_seed=RNG_SEED,
                   logit_scale=np.float64(float(model.logit_scale)), text_features=text.numpy(), logits0=logits0.numpy(),
                   pred_logits=pred.numpy(), plpd_threshold=np.float64(thr if thr is not None else args.plpd_threshold),
                   filter_ent=int(args.filter_ent), filter_plpd=int(args.filter_plpd), aug_type=str(args.aug_type))
        if thr is not None:
            rec["plpd"] = vals
        for i in spec.layers():
            for j, nm in enumerate(("A_q", "B_q", "A_v", "B_v")):
                p = lora_now[i][j]
                if nm.startswith("B"):
                    rec[f"lora_{i}_{nm}"] = p.detach().numpy().copy()
                    rec[f"grad_{i}_{nm}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        path = os.path.join(outdir, f"ref_b16_c10_deyo_{case}.npz")
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def _reference_plpd(deyo_ref, model, imgs, args, opt, optim_state, scaler):
    """Run the reference's forward_and_adapt_sar once with a keep-everything threshold and read back the PLPD vector it
    filtered on, by wrapping torch.where (the only consumer of `plpd > threshold`, deyo.py:146)."""
    import copy
    a = copy.copy(args)
    a.plpd_threshold = -10.0
    seen = {}
    orig = torch.where

    def spy(cond, *rest):
        seen.setdefault("conds", []).append(cond)
        return orig(cond, *rest)

    orig_gt = torch.Tensor.__gt__

    def gt(self, other):
        if isinstance(other, float) and other == -10.0:
            seen["plpd"] = self.detach().clone()
        return orig_gt(self, other)

    with torch.no_grad():
        model.LoRA_reset()
    opt.load_state_dict(optim_state)
    torch.manual_seed(RNG_SEED)
    torch.Tensor.__gt__ = gt
    try:
        deyo_ref.forward_and_adapt_sar(imgs, None, model, a, opt, scaler, a.deyo_margin, a.deyo_margin_e0)
    finally:
        torch.Tensor.__gt__ = orig_gt
    with torch.no_grad():
        model.LoRA_reset()
    return seen["plpd"].numpy()


if __name__ == "__main__":
    main()
