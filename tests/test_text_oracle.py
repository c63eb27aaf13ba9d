"""Pins oracle/text_oracle.py to the third-party code the reference actually runs for its class features:
transformers' CLIPTextModelWithProjection (clip/custom_clip.py:73-82 calls CLIPModel.get_text_features)."""
import numpy as np
import pytest
import torch

from oracle import text_oracle as TO


def test_text_oracle_matches_hf_clip_text_model():
    tr = pytest.importorskip("transformers")
    a = TO.TextArch(vocab=600, context=20, width=128, layers=3, heads=2, mlp=256, proj=48)
    cfg = tr.CLIPTextConfig(vocab_size=a.vocab, hidden_size=a.width, intermediate_size=a.mlp, projection_dim=a.proj,
                            num_hidden_layers=a.layers, num_attention_heads=a.heads, max_position_embeddings=a.context,
                            hidden_act="quick_gelu", layer_norm_eps=a.ln_eps, eos_token_id=2, bos_token_id=0, pad_token_id=1)
    torch.manual_seed(0)
    m = tr.CLIPTextModelWithProjection(cfg).eval()
    w = TO.make_synthetic_text_weights(a, seed=5)
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    tokens = TO.make_synthetic_tokens(7, a, seed=3)
    with torch.no_grad():
        ref = m(input_ids=tokens).text_embeds
    got = TO.text_forward(a, w, tokens, normalize=False)
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=0, atol=2e-5)
    unit = TO.text_forward(a, w, tokens)
    np.testing.assert_allclose(unit.norm(dim=-1).numpy(), np.ones(7), atol=1e-6)
