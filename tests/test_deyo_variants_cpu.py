"""Host logic of the optional DeYO branches (our deyo.py `_adapt_general` + `destroy_structure`)
against golden fixtures produced by the UNMODIFIED reference (oracle/make_golden_deyo_variants.py): `filter_ent`,
`filter_plpd` with the deterministic occlusion and with the seeded patch shuffle.  The model under the head is the fp32
CPU oracle (the checker), so the branch logic is compared at fp32 precision -- the PLPD values of a random-init model sit
within +-8e-3, far below bf16 noise, so the device path cannot be held to a threshold decision on this fixture."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ttl_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class OracleClip(nn.Module):
    """ClipTestTimeTuning-shaped wrapper over the oracle's functional forward (test infrastructure)."""

    def __init__(self, w, lora0, text, logit_scale):
        super().__init__()
        self.arch, self.spec, self.w, self.text, self.ls = O.ARCHS["ViT-B/16"], O.LoraSpec(), w, text, logit_scale
        self.layers = list(lora0)
        self.params = nn.ParameterList([nn.Parameter(t.clone()) for i in self.layers for t in lora0[i]])

    def lora(self):
        it = iter(self.params)
        return {i: [next(it) for _ in range(4)] for i in self.layers}

    def forward(self, x):
        return O.clip_logits(O.vision_forward(self.arch, self.w, x, self.lora(), self.spec.scale), self.text, self.ls)


@pytest.mark.slow
@pytest.mark.parametrize("case", ["fent", "plpd_occ", "plpd_patch"])
def test_deyo_branch_matches_reference(case, b16_weights, b16_views):
    import deyo
    g = np.load(os.path.join(GOLD, f"ref_b16_c10_deyo_{case}.npz"))
    spec = O.LoraSpec()
    lora0 = O.lora_init(O.ARCHS["ViT-B/16"], spec, seed=int(g["lora_seed"]))
    model = OracleClip(b16_weights, lora0, torch.from_numpy(g["text_features"]), float(g["logit_scale"]))
    opt = torch.optim.AdamW([{"params": [p]} for p in model.params], lr=5e-3)     # 12 groups like ttl.py:189-218
    scaler = torch.amp.GradScaler("cuda", init_scale=1000, enabled=False)
    args = types.SimpleNamespace(filter_ent=int(g["filter_ent"]), filter_plpd=int(g["filter_plpd"]), reweight_ent=1,
                                 reweight_plpd=0, selection_p=0.1, aug_type=str(g["aug_type"]), occlusion_size=112,
                                 patch_len=6, row_start=56, column_start=56, plpd_threshold=float(g["plpd_threshold"]))
    torch.manual_seed(int(g["rng_seed"]))
    out, backward, final = deyo.forward_and_adapt_sar(b16_views, None, model, args, opt, scaler, 0.5, 0.4)
    np.testing.assert_allclose(out.detach().numpy(), g["logits0"], rtol=0, atol=2e-4)
    if case == "fent":
        assert (backward, final) == (6, 6)
    else:
        assert backward == 64 and final == int((g["plpd"] > float(g["plpd_threshold"])).sum())
    lora = model.lora()
    for i in spec.layers():
        for j, nm in ((1, "B_q"), (3, "B_v")):
            ref_g, got_g = g[f"grad_{i}_{nm}"], lora[i][j].grad.numpy()
            assert np.linalg.norm(got_g - ref_g) / np.linalg.norm(ref_g) < 1e-4, (i, nm)
            big = np.abs(ref_g) > 0.1 * np.abs(ref_g).mean()       # step-1 Adam: p = -lr * g / (|g| + eps)
            np.testing.assert_allclose(lora[i][j].detach().numpy()[big], g[f"lora_{i}_{nm}"][big], rtol=0, atol=2e-5)
    with torch.no_grad():
        pred = model(b16_views[:1]).numpy()
    np.testing.assert_allclose(pred, g["pred_logits"], rtol=0, atol=5e-4)


def test_destroy_structure_shapes_and_invariants():
    import deyo
    x = torch.randn(3, 3, 224, 224)
    a = types.SimpleNamespace(aug_type="occ", occlusion_size=112, row_start=56, column_start=56, patch_len=6)
    y = deyo.destroy_structure(x, a)
    assert torch.equal(y[:, :, :56], x[:, :, :56]) and torch.allclose(y[:, :, 60, 60], x.reshape(3, 3, -1).mean(2))
    a.aug_type = "pixel"
    torch.manual_seed(0)
    y = deyo.destroy_structure(x, a)
    assert torch.allclose(y.reshape(3, 3, -1).sort(dim=2).values, x.reshape(3, 3, -1).sort(dim=2).values)
    a.aug_type = "patch"
    y = deyo.destroy_structure(x, a)
    assert y.shape == x.shape and not torch.equal(y, x)
    a.aug_type = "nope"
    with pytest.raises(ValueError):
        deyo.destroy_structure(x, a)
