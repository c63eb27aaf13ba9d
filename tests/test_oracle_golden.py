"""Pin oracle/ttl_oracle.py against outputs of the UNMODIFIED reference (tests/golden/ref_*.npz,
made by oracle/make_golden.py from /root/reference behind import shims).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import ttl_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("A_q", "B_q", "A_v", "B_v")


def _load(case):
    return np.load(os.path.join(GOLD, f"ref_b16_c10_{case}.npz"))


@pytest.mark.parametrize("case", ["tpt", "deyo", "tpt2"])
def test_oracle_reproduces_reference(case, b16_weights, b16_views):
    g = _load(case)
    arch, spec = O.ARCHS["ViT-B/16"], O.LoraSpec()
    assert int(g["weight_seed"]) == 1234 and int(g["image_seed"]) == 7
    lora0 = O.lora_init(arch, spec, seed=int(g["lora_seed"]))
    text = torch.from_numpy(g["text_features"])
    torch.set_num_threads(os.cpu_count() or 1)
    res = O.adapt_and_predict(arch, b16_weights, b16_views, text, float(g["logit_scale"]), lora0, spec,
                              head=str(g["head"]), tta_steps=int(g["tta_steps"]))
    # fp32 CPU vs fp32 CPU with the same op sequence: tight tolerance (observed 0.0 in the dev container)
    np.testing.assert_allclose(res.logits0.numpy(), g["logits0"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(res.entropies.numpy(), g["entropies"], rtol=0, atol=2e-6)
    if str(g["head"]) == "tpt":
        assert sorted(res.idx.tolist()) == g["idx_sorted"].tolist()      # selection indices: bit-exact
    np.testing.assert_allclose(res.pred_logits.numpy(), g["pred_logits"], rtol=0, atol=2e-5)
    for i in spec.layers():
        for j, nm in enumerate(NAMES):
            ref_g = g[f"grad_{i}_{nm}"]
            got_g = res.grads[i][j].numpy()
            denom = max(np.linalg.norm(ref_g), 1e-30)
            assert np.linalg.norm(got_g - ref_g) / denom < 1e-4 or np.abs(ref_g).max() == 0, (i, nm)
            np.testing.assert_allclose(res.lora[i][j].numpy(), g[f"lora_{i}_{nm}"], rtol=0, atol=1e-6)


def test_golden_structural_facts():
    """at step 1 dA == 0 exactly and B == -lr*g/(|g|+eps); at step 2 dA != 0."""
    g = _load("tpt")
    for i in (9, 10, 11):
        assert np.abs(g[f"grad_{i}_A_q"]).max() == 0 and np.abs(g[f"grad_{i}_A_v"]).max() == 0
        gb = g[f"grad_{i}_B_v"]
        np.testing.assert_allclose(g[f"lora_{i}_B_v"], -5e-3 * gb / (np.abs(gb) + 1e-8), atol=2e-8)
    g2 = _load("tpt2")
    assert np.abs(g2["grad_10_A_q"]).max() > 0
