"""Parity of the CUDA path at the configurations that are actually benched (BASELINE.json configs[1..4]), through the C ABI:

  config 2  ViT-B/16, 1000 ImageNet class prompts, 64 views, S = 9 concurrent samples (bench.py / CLI default: M = 113 472
            rows, 576-column adapter pack): the reference's own fixture rides in one slot, the live CPU oracle checks
            another, all nine slots equal consecutive single-sample calls.
  config 3  the 200 ImageNet-A class prompts, both heads, bf16.
  config 4  ViT-L/14 @224, layers 21-23, 64 views, against the oracle fixture.
  config 5  4 TTA steps (per-step losses), 2 x 2 DeYO steps.

Fixtures: oracle/make_golden_configs.py (unmodified reference behind oracle/ref_shim.py; ViT-L/14 from the pinned oracle,
which the reference cannot build).  Tolerances are BASELINE.json's: logits and LoRA gradients within 1e-2 relative
(norm-wise) in bf16 with the selected views teacher-forced; adapted top-1 identical.  Where a later optimiser step sits on
sign-like earlier updates the gradient bound is stated next to the measured sign-flip fraction."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ttl_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("A_q", "B_q", "A_v", "B_v")
PRED_TOL = 2e-2     # one view, bf16: see tests/test_gpu_e2e.py


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _check_step1_grads(eng, g, layers, sample=0, tol=1e-2):
    from ttl_b200 import _lib as L
    worst = 0.0
    for i in layers:
        for j, nm in enumerate(NAMES):
            ref_g, got_g = g[f"grad_{i}_{nm}"], eng.lora_get(i, j, L.LORA_GRAD, sample=sample)
            if nm.startswith("A"):
                assert np.abs(got_g).max() == 0.0 and np.abs(ref_g).max() == 0.0           # dA == 0 exactly at step 1
                continue
            worst = max(worst, _rel(got_g, ref_g))
            got_p, ref_p = eng.lora_get(i, j, sample=sample), g[f"lora_{i}_{nm}"]
            mask = np.abs(ref_g) > 0.1 * np.abs(ref_g).mean()
            assert _rel(got_p[mask], ref_p[mask]) < 1e-2, (i, nm)
    assert worst < tol, worst
    return worst


def test_config2_nine_concurrent_samples_c1000(b16_weights, b16_views):
    """The benched shape.  Slot 2 = the reference's 1000-class fixture sample, slot 6 = a sample checked by the live oracle
    (~6 s of CPU), the other seven are further synthetic samples; forced views for slots 2 and 6, free-running elsewhere is
    not possible in one call, so the free-running selection of a first call is fed back as the forced set."""
    from ttl_b200 import Engine, Hparams
    from ttl_b200 import _lib as L
    g = np.load(os.path.join(GOLD, "ref_b16_c1000_tpt.npz"))
    arch, spec = O.ARCHS["ViT-B/16"], O.LoraSpec()
    lora0 = O.lora_init(arch, spec, seed=0)
    text = torch.from_numpy(g["text_features"])
    scale = float(g["logit_scale"])
    S, V = 9, 64
    imgs = [O.make_synthetic_views(V, arch.image_size, seed=400 + i) for i in range(S)]
    imgs[2] = b16_views
    torch.set_num_threads(os.cpu_count() or 1)
    ref6 = O.adapt_and_predict(arch, b16_weights, imgs[6], text, scale, lora0, spec, head="tpt")
    eng = Engine("ViT-B/16", max_views=V, max_classes=1000, layer_range=(9, 11), max_samples=S)
    try:
        eng.load_weights(b16_weights)
        eng.set_lora_init(lora0)
        eng.set_text_features(text, scale)
        hp = Hparams(head="tpt")
        batch = torch.stack(imgs).cuda()
        want = ("logits0", "entropy", "idx", "loss", "pred_logits")
        free = eng.adapt_predict_batch(batch, hp, want=want)
        forced = free["idx"].clone()
        forced[2] = torch.from_numpy(g["idx"].astype(np.int32))
        forced[6] = ref6.idx.to(torch.int32)
        for rep in range(3):                       # eager, capture, replay
            out = {k: v.cpu() for k, v in eng.adapt_predict_batch(batch, hp, forced_idx=forced, want=want).items()}
        # slot 2 against the unmodified reference
        assert _rel(out["logits0"][2].numpy(), g["logits0"]) < 1e-2
        assert float(np.abs(out["entropy"][2].numpy() - g["entropies"]).max()) < 2e-2
        assert abs(float(out["loss"][2]) - float(g["losses"][0])) < 1e-2 * max(1.0, abs(float(g["losses"][0])))
        assert _rel(out["pred_logits"][2].numpy(), g["pred_logits"][0]) < PRED_TOL
        assert int(out["pred_logits"][2].argmax()) == int(g["pred_logits"][0].argmax())
        w2 = _check_step1_grads(eng, g, spec.layers(), sample=2)
        # slot 6 against the live oracle
        assert _rel(out["logits0"][6].numpy(), ref6.logits0.numpy()) < 1e-2
        assert abs(float(out["loss"][6]) - ref6.loss) < 1e-2 * max(1.0, abs(ref6.loss))
        assert _rel(out["pred_logits"][6].numpy(), ref6.pred_logits[0].numpy()) < PRED_TOL
        assert int(out["pred_logits"][6].argmax()) == int(ref6.pred_logits[0].argmax())
        w6 = 0.0
        for i in spec.layers():
            for j in (1, 3):
                w6 = max(w6, _rel(eng.lora_get(i, j, L.LORA_GRAD, sample=6), ref6.grads[i][j].numpy()))
        assert w6 < 1e-2, w6
        print(f"[config 2, S=9, C=1000] worst dB rel err: reference slot {w2:.2e}, live-oracle slot {w6:.2e}")
        # all nine slots against consecutive single-sample calls
        eng.set_graphs(False)
        for sidx in range(S):
            one = eng.adapt_predict(batch[sidx], hp, forced_idx=forced[sidx], want=want)
            assert _rel(out["logits0"][sidx].numpy(), one["logits0"].cpu().numpy()) < 1e-5, sidx
            assert out["idx"][sidx].tolist() == one["idx"].cpu().tolist(), sidx
            assert abs(float(out["loss"][sidx]) - float(one["loss"])) < 1e-4 * max(1.0, abs(float(one["loss"]))), sidx
            # same kernels per sample; the bound covers tile-shape dependent rounding at M = 113 472 vs 12 608 and the few step-1 sign
            # flips of near-zero gradient elements that follow from it (measured up to 2.4e-3)
            assert _rel(out["pred_logits"][sidx].numpy(), one["pred_logits"].cpu().numpy()) < 4e-3, sidx
            assert int(out["pred_logits"][sidx].argmax()) == int(one["pred_logits"].argmax()), sidx
    finally:
        eng.close()


@pytest.mark.parametrize("case", ["ref_b16_c1000_tpt", "ref_b16_c200_tpt", "ref_b16_c200_deyo"])
def test_config2_config3_fixtures_bf16(b16_weights, b16_views, case):
    """1000 / 200 classes through the reference's own prompt builder and text tower, one sample per call, both heads."""
    from ttl_b200 import Engine, Hparams
    g = np.load(os.path.join(GOLD, case + ".npz"))
    arch, spec = O.ARCHS["ViT-B/16"], O.LoraSpec()
    eng = Engine("ViT-B/16", max_views=64, max_classes=1000, layer_range=(9, 11))
    try:
        eng.load_weights(b16_weights)
        eng.set_lora_init(O.lora_init(arch, spec, seed=0))
        eng.set_text_features(g["text_features"], float(g["logit_scale"]))
        head = str(g["head"])
        forced = torch.from_numpy(g["idx"].astype(np.int32)) if head == "tpt" else None
        for rep in range(3):
            out = eng.adapt_predict(b16_views.cuda(), Hparams(head=head), forced_idx=forced,
                                    want=("logits0", "entropy", "idx", "loss", "pred_logits"))
        assert _rel(out["logits0"].cpu().numpy(), g["logits0"]) < 1e-2
        assert float(np.abs(out["entropy"].cpu().numpy() - g["entropies"]).max()) < 2e-2
        assert _rel(out["pred_logits"].cpu().numpy(), g["pred_logits"][0]) < PRED_TOL
        assert int(out["pred_logits"].argmax()) == int(g["pred_logits"][0].argmax())
        if head == "tpt":
            assert out["idx"].cpu().tolist() == g["idx"].tolist()
            assert abs(float(out["loss"]) - float(g["losses"][0])) < 1e-2 * max(1.0, abs(float(g["losses"][0])))
        w = _check_step1_grads(eng, g, spec.layers())
        print(f"[{case}] worst dB rel err {w:.2e}")
    finally:
        eng.close()


def test_config5_four_steps_vs_reference(b16_weights, b16_views):
    """tta_steps = 4 (ttl.py:90-108; selected_idx frozen after step 1).  Per-step losses: the library returns the loss of the
    last optimiser step, so steps 1..4 are four calls.  Loss and adapted prediction are held to 1e-2 / PRED_TOL.  Gradients of
    steps >= 2 are taken at factors whose first update was -lr * sign(g): the elements whose sign differs between bf16 and fp32
    (measured and printed below; 0.2-0.6 % per SURVEY.md 7.3-1) move the operating point by 2 * lr each, so the bound on the
    step-4 gradient is looser than the single-step 1e-2 and is tied to that fraction."""
    from ttl_b200 import Engine, Hparams
    from ttl_b200 import _lib as L
    g = np.load(os.path.join(GOLD, "ref_b16_c10_tpt4.npz"))
    g1 = np.load(os.path.join(GOLD, "ref_b16_c10_tpt.npz"))
    arch, spec = O.ARCHS["ViT-B/16"], O.LoraSpec()
    eng = Engine("ViT-B/16", max_views=64, max_classes=16, layer_range=(9, 11))
    try:
        eng.load_weights(b16_weights)
        eng.set_lora_init(O.lora_init(arch, spec, seed=0))
        eng.set_text_features(g["text_features"], float(g["logit_scale"]))
        forced = torch.from_numpy(g["idx"].astype(np.int32))
        imgs = b16_views.cuda()
        losses = []
        for steps in (1, 2, 3, 4):
            out = eng.adapt_predict(imgs, Hparams(head="tpt", tta_steps=steps), forced_idx=forced, want=("loss", "pred_logits"))
            losses.append(float(out["loss"]))
            if steps == 1:      # sign flips of the first update, against the single-step fixture
                flips = []
                for i in spec.layers():
                    for j in (1, 3):
                        flips.append(float((np.sign(eng.lora_get(i, j)) != np.sign(g1[f"lora_{i}_{NAMES[j]}"])).mean()))
        ref_losses = g["losses"]
        print(f"[config 5] per-step losses {losses} vs reference {ref_losses.tolist()}; first-update sign flips "
              f"{min(flips):.4f}..{max(flips):.4f}")
        for got, ref in zip(losses, ref_losses):
            assert abs(got - ref) < 1e-2 * max(1.0, abs(ref)), (losses, ref_losses)
        assert _rel(out["pred_logits"].cpu().numpy(), g["pred_logits"][0]) < PRED_TOL
        assert int(out["pred_logits"].argmax()) == int(g["pred_logits"][0].argmax())
        worst = 0.0
        for i in spec.layers():
            for j, nm in enumerate(NAMES):
                worst = max(worst, _rel(eng.lora_get(i, j, L.LORA_GRAD), g[f"grad_{i}_{nm}"]))
        print(f"[config 5] worst step-4 LoRA-gradient rel err {worst:.3e}")
        # 2 * sqrt(flip fraction) is the norm-wise distance of two +-lr sign patterns ('biggest parity risk')
        assert worst < max(1e-2, 4.0 * math.sqrt(max(flips))), (worst, flips)
    finally:
        eng.close()


@pytest.mark.parametrize("point", ["step2", "step4"])
def test_config5_later_step_gradient_at_the_reference_operating_point(b16_weights, b16_views, point):
    """Gradient of a LATER optimiser step with the factors the REFERENCE had before that step loaded as the starting point:
    B != 0, so the LoRA branch of the forward, U = dY B, dA and dB all run, and -- the operating point being shared -- they can
    be held to the single-step bar instead of the sign-flip bound of the free-running test above.
      step2: factors after the reference's first step (`lora_*` of the 1-step fixture) -> `grad_*` of the 2-step fixture.
      step4: `prelast_*` of the 4-step fixture (captured by wrapping the reference optimiser's step) -> its `grad_*`.
    fp32 mode: 1e-4 at both points.  bf16 path: 1e-2 at step 2; at step 4 the loss has fallen to 0.016 (the marginal distribution
    is nearly one-hot), dlogits is a difference of nearly cancelling softmax terms and the 1.6e-3 bf16 noise of the logits shows
    up amplified: bound 3e-2, measured value printed."""
    from ttl_b200 import Engine, Hparams
    from ttl_b200 import _lib as L
    spec = O.LoraSpec()
    if point == "step2":
        g0, g = np.load(os.path.join(GOLD, "ref_b16_c10_tpt.npz")), np.load(os.path.join(GOLD, "ref_b16_c10_tpt2.npz"))
        assert g0["idx_sorted"].tolist() == g["idx_sorted"].tolist()
        start = {i: [torch.from_numpy(g0[f"lora_{i}_{nm}"]) for nm in NAMES] for i in spec.layers()}
        bf16_tol = 1e-2
    else:
        g = np.load(os.path.join(GOLD, "ref_b16_c10_tpt4.npz"))
        start = {i: [torch.from_numpy(g[f"prelast_{i}_{nm}"]) for nm in NAMES] for i in spec.layers()}
        bf16_tol = 3e-2
    assert min(float(start[i][1].abs().max()) for i in spec.layers()) > 0.0          # B != 0 at this point
    forced = torch.from_numpy(g["idx_sorted"].astype(np.int32))      # the loss is a mean over the selected views: order-free
    ref_loss = float(g["losses"][-1]) if "losses" in g.files and len(g["losses"]) else None
    for precision, tol in (("fp32", 1e-4), ("bf16", bf16_tol)):
        eng = Engine("ViT-B/16", max_views=64, max_classes=16, layer_range=(9, 11), precision=precision)
        try:
            eng.load_weights(b16_weights)
            eng.set_lora_init(start)
            eng.set_text_features(g["text_features"], float(g["logit_scale"]))
            out = eng.adapt_predict(b16_views.cuda(), Hparams(head="tpt", tta_steps=1), forced_idx=forced, want=("loss",))
            worst = {"A": 0.0, "B": 0.0}
            for i in spec.layers():
                for j, nm in enumerate(NAMES):
                    assert np.abs(g[f"grad_{i}_{nm}"]).max() > 0
                    worst[nm[0]] = max(worst[nm[0]], _rel(eng.lora_get(i, j, L.LORA_GRAD), g[f"grad_{i}_{nm}"]))
            print(f"[config 5, {point} gradient at the reference's factors, {precision}] dA {worst['A']:.3e}, dB {worst['B']:.3e}, "
                  f"loss {float(out['loss']):.6f} (reference {ref_loss})")
            assert worst["A"] < tol and worst["B"] < tol, (precision, worst)
            if ref_loss is not None:
                assert abs(float(out["loss"]) - ref_loss) < max(tol, 1e-4) * max(1.0, abs(ref_loss))
        finally:
            eng.close()


def test_config5_concurrent_four_steps_equal_sequential(b16_weights):
    """config 5 as benched: S = 9 samples x 4 steps in one call == nine consecutive single-sample calls."""
    from ttl_b200 import Engine, Hparams
    arch = O.ARCHS["ViT-B/16"]
    S, V = 9, 64
    eng = Engine("ViT-B/16", max_views=V, max_classes=256, layer_range=(9, 11), max_samples=S)
    try:
        eng.load_weights(b16_weights)
        eng.set_lora_init(O.lora_init(arch, O.LoraSpec(), seed=0))
        eng.set_text_features(O.make_text_features(200, arch.proj), math.log(100.0))
        hp = Hparams(head="tpt", tta_steps=4)
        imgs = torch.stack([O.make_synthetic_views(V, arch.image_size, seed=700 + i) for i in range(S)]).cuda()
        want = ("idx", "loss", "pred_logits")
        for rep in range(3):
            batch = {k: v.cpu() for k, v in eng.adapt_predict_batch(imgs, hp, want=want).items()}
        eng.set_graphs(False)
        for s in range(S):
            one = eng.adapt_predict(imgs[s], hp, want=want)
            assert batch["idx"][s].tolist() == one["idx"].cpu().tolist(), s
            # four sign-like updates deep: tile-shape dependent rounding of the big-M vs small-M kernels shows at 1e-2
            assert abs(float(batch["loss"][s]) - float(one["loss"])) < 2e-2 * max(1.0, abs(float(one["loss"]))), s
            assert _rel(batch["pred_logits"][s].numpy(), one["pred_logits"].cpu().numpy()) < PRED_TOL, s
            assert int(batch["pred_logits"][s].argmax()) == int(one["pred_logits"].argmax()), s
    finally:
        eng.close()


def test_config5_deyo_two_by_two_steps_vs_reference(b16_weights, b16_views):
    """tta_steps = 2 under the script-default head = 4 optimiser steps."""
    from ttl_b200 import Engine, Hparams
    g = np.load(os.path.join(GOLD, "ref_b16_c10_deyo2.npz"))
    arch, spec = O.ARCHS["ViT-B/16"], O.LoraSpec()
    eng = Engine("ViT-B/16", max_views=64, max_classes=16, layer_range=(9, 11))
    try:
        eng.load_weights(b16_weights)
        eng.set_lora_init(O.lora_init(arch, spec, seed=0))
        eng.set_text_features(g["text_features"], float(g["logit_scale"]))
        out = eng.adapt_predict(b16_views.cuda(), Hparams(head="deyo", tta_steps=2), want=("logits0", "pred_logits"))
        assert _rel(out["logits0"].cpu().numpy(), g["logits0"]) < 1e-2
        e = _rel(out["pred_logits"].cpu().numpy(), g["pred_logits"][0])
        print(f"[config 5, DeYO 2x2] adapted prediction rel err {e:.3e}")
        assert e < 3e-2 and int(out["pred_logits"].argmax()) == int(g["pred_logits"][0].argmax())
    finally:
        eng.close()


def test_config4_vit_l14_64_views_vs_oracle():
    """ViT-L/14 @224 (257 tokens, d = 1024, 24 layers, adapters on 21-23), 64 views, against the oracle fixture."""
    from ttl_b200 import Engine, Hparams
    from ttl_b200 import _lib as L
    g = np.load(os.path.join(GOLD, "oracle_l14_c10_tpt.npz"))
    arch = O.ARCHS["ViT-L/14"]
    spec = O.LoraSpec(rank=16, alpha=32.0, layer_lo=21, layer_hi=23)
    w = O.make_synthetic_weights(arch, int(g["weight_seed"]))
    imgs = O.make_synthetic_views(64, arch.image_size, seed=int(g["image_seed"]))
    eng = Engine("ViT-L/14", max_views=64, max_classes=16, layer_range=(21, 23))
    try:
        eng.load_weights(w)
        eng.set_lora_init(O.lora_init(arch, spec, int(g["lora_seed"])))
        eng.set_text_features(g["text_features"], float(g["logit_scale"]))
        forced = torch.from_numpy(g["idx"].astype(np.int32))
        for rep in range(3):
            out = eng.adapt_predict(imgs.cuda(), Hparams(head="tpt"), forced_idx=forced,
                                    want=("logits0", "entropy", "loss", "pred_logits"))
        e_log = _rel(out["logits0"].cpu().numpy(), g["logits0"])
        e_pred = _rel(out["pred_logits"].cpu().numpy(), g["pred_logits"][0])
        worst = 0.0
        for i in spec.layers():
            for j in (1, 3):
                worst = max(worst, _rel(eng.lora_get(i, j, L.LORA_GRAD), g[f"grad_{i}_{NAMES[j]}"]))
        print(f"[config 4, ViT-L/14, 64 views] logits {e_log:.3e}, dB {worst:.3e}, adapted prediction {e_pred:.3e}")
        assert e_log < 1e-2
        assert abs(float(out["loss"]) - float(g["losses"][0])) < 1e-2 * max(1.0, abs(float(g["losses"][0])))
        assert worst < 1e-2
        assert e_pred < PRED_TOL and int(out["pred_logits"].argmax()) == int(g["pred_logits"][0].argmax())
    finally:
        eng.close()
