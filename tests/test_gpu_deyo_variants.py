"""The optional branches of the reference's weighted-entropy head on the fused CUDA path (reference
deyo.py:103-151): `filter_ent`, `filter_plpd` with `--aug_type occ / patch / pixel`, the re-weighting switches, and a sample
that keeps no view.  Fixtures: the unmodified reference's forward_and_adapt_sar (oracle/make_golden_deyo_variants.py; the
branch logic itself is pinned to the reference on CPU in tests/test_deyo_variants_cpu.py)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ttl_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PRED_TOL = 2e-2


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="module")
def eng(b16_weights):
    from ttl_b200 import Engine
    e = Engine("ViT-B/16", max_views=64, max_classes=16, layer_range=(9, 11), max_samples=3)
    e.load_weights(b16_weights)
    e.set_lora_init(O.lora_init(O.ARCHS["ViT-B/16"], O.LoraSpec(), seed=0))
    yield e
    e.close()


@pytest.mark.parametrize("case", ["fent", "plpd_occ", "plpd_patch"])
def test_fused_optional_branches_vs_reference(eng, b16_views, case):
    from ttl_b200 import Hparams
    from ttl_b200 import _lib as L
    g = np.load(os.path.join(GOLD, f"ref_b16_c10_deyo_{case}.npz"))
    eng.set_text_features(g["text_features"], float(g["logit_scale"]))
    kw = dict(filter_ent=int(g["filter_ent"]), filter_plpd=int(g["filter_plpd"]), plpd_threshold=float(g["plpd_threshold"]),
              aug_type=str(g["aug_type"]))
    # The PLPD values of this random-init model span +-8e-3 and the fixture's threshold sits in a 4e-4 gap: narrower than the bf16
    # noise on a probability (measured 4-5e-4), so one view can cross it.  As for the selection (forced_idx), the reference's own
    # filter outcome is teacher-forced for the gradient comparison, and the free-running values are held to the noise band.
    forced = (g["plpd"] > float(g["plpd_threshold"])).astype(np.int32)[None] if kw["filter_plpd"] else None
    torch.manual_seed(int(g["rng_seed"]))          # the tile orders of aug_type patch come from the CPU torch RNG, as in the reference
    out = eng.adapt_predict_batch_deyo(b16_views.cuda()[None], Hparams(head="deyo"), want=("logits0", "pred_logits", "idx", "loss"),
                                       forced_keep=forced, **kw)
    torch.cuda.synchronize()
    assert _rel(out["logits0"][0].cpu().numpy(), g["logits0"]) < 1e-2
    plpd, n_final = eng.deyo_last_plpd()
    if kw["filter_ent"]:
        ent = O.softmax_entropy(torch.from_numpy(g["logits0"]))
        want_idx = torch.argsort(ent, stable=True)[:6].tolist()
        assert sorted(out["idx"][0].cpu().tolist()) == sorted(want_idx) and int(n_final[0]) == 6
    if kw["filter_plpd"]:
        ref_plpd, thr = g["plpd"], kw["plpd_threshold"]
        dev = float(np.abs(plpd[0] - ref_plpd).max())
        flipped = np.nonzero((plpd[0] > thr) != (ref_plpd > thr))[0]
        print(f"[deyo {case}] PLPD max abs dev {dev:.2e} (values span {ref_plpd.min():.2e}..{ref_plpd.max():.2e}, threshold gap "
              f"{np.abs(ref_plpd - thr).min():.2e}), views across the threshold: {flipped.tolist()}")
        assert dev < 1e-3
        assert int(n_final[0]) == int(forced.sum())
        assert all(abs(float(ref_plpd[v]) - thr) <= dev for v in flipped)      # only views inside the noise band may cross
    worst = 0.0
    for i in (9, 10, 11):
        for j, nm in ((1, "B_q"), (3, "B_v")):
            worst = max(worst, _rel(eng.lora_get(i, j, L.LORA_GRAD), g[f"grad_{i}_{nm}"]))
    print(f"[deyo {case}] worst dB rel err {worst:.2e}, kept {int(n_final[0])}")
    assert worst < 1e-2
    assert _rel(out["pred_logits"][0].cpu().numpy(), g["pred_logits"][0]) < PRED_TOL
    assert int(out["pred_logits"][0].argmax()) == int(g["pred_logits"][0].argmax())


def test_destroyed_views_match_the_reference_transform(eng, b16_views):
    """x' of deyo.py:116-136 for the three --aug_type values against torch / torchvision run on the same tensors: the kernels are
    reached through a PLPD threshold that keeps everything, so compare the PLPD values against a second library forward on the
    torch-built x' (same model, same draws)."""
    import deyo as deyo_mod
    import types
    from ttl_b200 import Hparams
    eng.set_text_features(O.make_text_features(10, 512, seed=3), math.log(100.0))
    imgs = b16_views[:16].cuda()
    for aug in ("occ", "patch", "pixel"):
        args = types.SimpleNamespace(aug_type=aug, occlusion_size=112, row_start=56, column_start=56, patch_len=6)
        torch.manual_seed(77)
        xp = deyo_mod.destroy_structure(imgs, args)                 # torch ops + the CPU RNG, as the reference
        eng.lora_reset()
        lg, lgp = eng.forward(imgs), eng.forward(xp)
        p, pp = lg.softmax(1), lgp.softmax(1)
        top = p.argmax(1, keepdim=True)
        want = (p.gather(1, top) - pp.gather(1, top)).flatten().cpu().numpy()
        torch.manual_seed(77)
        eng.adapt_predict_batch_deyo(imgs[None], Hparams(head="deyo"), filter_plpd=1, plpd_threshold=-10.0, aug_type=aug)
        got, n_final = eng.deyo_last_plpd()
        assert int(n_final[0]) == 16
        assert float(np.abs(got[0] - want).max()) < 5e-3, (aug, float(np.abs(got[0] - want).max()))


def test_concurrent_samples_and_a_sample_without_kept_views(eng):
    """Three samples in one call == three single-sample calls (own draws per sample, in the reference's order); and a PLPD
    threshold nothing passes: no optimiser step (deyo.py:184), the prediction is the un-adapted model's."""
    from ttl_b200 import Hparams
    from ttl_b200 import _lib as L
    arch = O.ARCHS["ViT-B/16"]
    eng.set_text_features(O.make_text_features(10, 512, seed=3), math.log(100.0))
    imgs = torch.stack([O.make_synthetic_views(64, arch.image_size, seed=900 + i) for i in range(3)]).cuda()
    hp = Hparams(head="deyo")
    kw = dict(filter_ent=1, filter_plpd=1, plpd_threshold=-10.0, aug_type="patch")
    torch.manual_seed(5)
    perm = eng.draw_deyo_perms(3, 1, 6, "patch", 6, 224)
    batch = eng.adapt_predict_batch_deyo(imgs, hp, perm=perm, want=("pred_logits", "idx", "loss"), **kw)
    for s in range(3):
        one = eng.adapt_predict_batch_deyo(imgs[s:s + 1], hp, perm=perm[s:s + 1], want=("pred_logits", "idx", "loss"), **kw)
        assert batch["idx"][s].tolist() == one["idx"][0].tolist()
        assert abs(float(batch["loss"][s]) - float(one["loss"][0])) < 1e-4 * max(1.0, abs(float(one["loss"][0])))
        assert _rel(batch["pred_logits"][s].cpu().numpy(), one["pred_logits"][0].cpu().numpy()) < 2e-3
    # nothing survives the filter
    out = eng.adapt_predict_batch_deyo(imgs[:1], hp, filter_plpd=1, plpd_threshold=10.0, aug_type="occ", want=("pred_logits",))
    _, n_final = eng.deyo_last_plpd()
    assert int(n_final[0]) == 0
    for i in (9, 10, 11):
        assert np.abs(eng.lora_get(i, 1) - 0.0).max() == 0.0 and np.abs(eng.lora_get(i, 3)).max() == 0.0     # B still at its reset value
    plain = eng.adapt_predict_batch(imgs[:1], Hparams(head="deyo", tta_steps=0))["pred_logits"]     # same kernels, no step
    assert _rel(out["pred_logits"][0].cpu().numpy(), plain[0].cpu().numpy()) < 1e-6


def test_reweight_switches(eng, b16_views):
    """reweight_ent = 0 and reweight_plpd = 0: plain mean entropy (deyo.py:159: the coefficient branch is skipped)."""
    from ttl_b200 import Hparams
    eng.set_text_features(O.make_text_features(10, 512, seed=3), math.log(100.0))
    imgs = b16_views.cuda()[None]
    out = eng.adapt_predict_batch_deyo(imgs, Hparams(head="deyo"), reweight_ent=0, reweight_plpd=0, want=("logits0", "loss"))
    ent = O.softmax_entropy(out["logits0"][0].cpu().double())
    assert abs(float(out["loss"][0]) - float(ent.mean())) < 1e-4
    out = eng.adapt_predict_batch_deyo(imgs, Hparams(head="deyo"), reweight_ent=1, reweight_plpd=1, want=("logits0", "loss"))
    w = torch.exp(-(ent - 0.4))
    assert abs(float(out["loss"][0]) - float((w * ent).mean())) < 1e-4
