"""Pin oracle/text_oracle.py:adapt_and_predict_text_lora -- the checker of the `--lora_encoder text` variant (SURVEY.md 8f N4:
adapter on the text tower, class features recomputed with gradient in every forward) -- against outputs of the UNMODIFIED
reference (tests/golden/ref_b16_c10_textlora_*.npz, made by oracle/make_golden_text_lora.py behind the import shims).
CPU only.  The CUDA library does not implement this variant yet; the oracle comes first."""
import os

import numpy as np
import pytest
import torch

from oracle import text_oracle as TO
from oracle import ttl_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("A_q", "B_q", "A_v", "B_v")


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="module")
def towers(b16_weights):
    tarch = TO.TEXT_ARCHS["ViT-B/16"]
    return O.ARCHS["ViT-B/16"], b16_weights, tarch, TO.make_synthetic_text_weights(tarch, 4321)


@pytest.mark.parametrize("head", ["tpt", "deyo"])
def test_text_lora_oracle_reproduces_reference(head, towers, b16_views):
    varch, vw, tarch, tw = towers
    g = np.load(os.path.join(GOLD, f"ref_b16_c10_textlora_{head}.npz"))
    assert int(g["weight_seed"]) == 1234 and int(g["text_weight_seed"]) == 4321 and int(g["image_seed"]) == 7
    tokens = torch.from_numpy(g["tokens"])
    lora0 = TO.text_lora_init(tarch, range(9, 12), seed=int(g["lora_seed"]))
    r = TO.adapt_and_predict_text_lora(varch, vw, tarch, tw, tokens, b16_views, float(g["logit_scale"]), lora0, head=head)
    assert _rel(r.logits0.numpy(), g["logits0"]) < 1e-5
    assert _rel(r.entropies.numpy(), g["entropies"]) < 1e-5
    if head == "tpt":
        assert np.array_equal(np.sort(r.idx.numpy()), g["idx_sorted"])         # integer indices: exact
    for i in range(9, 12):
        for j, nm in enumerate(NAMES):
            gr = g[f"grad_{i}_{nm}"]
            if nm.startswith("A"):
                assert not gr.any() and not r.grads[i][j].numpy().any()        # B = 0 at step 1  =>  dA == 0 exactly
                assert _rel(r.lora[i][j].numpy(), g[f"lora_{i}_{nm}"]) < 1e-6  # only the weight decay moves A
            else:
                assert _rel(r.grads[i][j].numpy(), gr) < 1e-4
                big = np.abs(gr) > 0.1 * np.abs(gr).mean()                     # step-1 Adam is -lr * sign(g): compare where |g| is not tiny
                assert _rel(r.lora[i][j].numpy()[big], g[f"lora_{i}_{nm}"][big]) < 1e-3
    assert _rel(r.pred_logits.numpy(), g["pred_logits"]) < 1e-4


def test_text_lora_is_inert_while_b_is_zero_and_only_in_range_layers_matter():
    """B = 0 => the LoRA branch contributes exactly nothing; an entry for a layer contributes only through that layer."""
    a = TO.TEXT_ARCHS["tiny"]
    w = TO.make_synthetic_text_weights(a, 5)
    tok = TO.make_synthetic_tokens(6, a, 3)
    base = TO.text_forward(a, w, tok)
    lora = TO.text_lora_init(a, range(0, 2), seed=1)
    assert torch.equal(TO.text_forward(a, w, tok, True, lora), base)
    lora[1][1] = torch.randn(a.width, 16) * 0.05
    moved = TO.text_forward(a, w, tok, True, lora)
    assert (moved - base).abs().max() > 1e-4
    assert torch.equal(TO.text_forward(a, w, tok, True, {0: lora[0]}), base)     # layer 0 still has B = 0
