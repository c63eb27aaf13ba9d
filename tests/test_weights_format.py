"""OpenAI-format checkpoint mapping (ttl_b200/weights.py) pinned to the reference's own OpenAI-format
model: tests/golden/openai_vit_tiny.npz holds the state dict of clip/model.py `VisionTransformer` (tiny geometry, seeded),
seeded images and ITS output features (oracle/make_golden_openai_format.py).  Converted to HF names and pushed through the
oracle's HF-style encoder restatement, the features must agree at fp32 precision -- which pins the name mapping, the
in_proj q/k/v split, the `x @ proj` transpose and, once more, the oracle's encoder math."""
import os

import numpy as np
import pytest
import torch

from oracle import ttl_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "openai_vit_tiny.npz")


def _load():
    g = np.load(GOLD)
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    return g, sd


def test_openai_state_dict_maps_onto_the_hf_layout():
    from ttl_b200.weights import is_openai_format, openai_to_hf_vision
    g, sd = _load()
    assert is_openai_format(sd)
    hf = openai_to_hf_vision(sd)
    arch = O.VitArch("tiny-openai", int(g["geom_input_resolution"]), int(g["geom_patch_size"]), int(g["geom_width"]),
                     int(g["geom_layers"]), int(g["geom_heads"]), 4 * int(g["geom_width"]), int(g["geom_output_dim"]))
    feats = O.vision_forward(arch, hf, torch.from_numpy(g["images"]))
    np.testing.assert_allclose(feats.numpy(), g["features"], rtol=0, atol=2e-5)
    # every tensor of the vision tower found a slot, nothing else leaked in
    assert len(hf) == 8 + 16 * arch.layers
    assert hf["visual_projection.weight"].shape == (arch.proj, arch.width)


def test_hf_state_dicts_pass_through_and_foreign_ones_are_rejected():
    from ttl_b200.weights import hf_vision_subset, is_openai_format, openai_to_hf_vision
    w = O.make_synthetic_weights(O.ARCHS["ViT-tiny"], 1)
    w["text_model.embeddings.token_embedding.weight"] = torch.zeros(4, 4)
    sub = hf_vision_subset(w)
    assert not is_openai_format(w) and "text_model.embeddings.token_embedding.weight" not in sub
    assert set(sub) == {k for k in w if k.startswith("vision_model.") or k == "visual_projection.weight"}
    with pytest.raises(KeyError):
        openai_to_hf_vision({"something.else": torch.zeros(1)})


def test_checkpoint_files_round_trip(tmp_path):
    from ttl_b200.weights import load_vision_checkpoint
    g, sd = _load()
    p = tmp_path / "ViT-tiny.pt"
    torch.save(sd, p)
    a = load_vision_checkpoint(str(p))
    assert "vision_model.encoder.layers.1.self_attn.v_proj.bias" in a
    st = pytest.importorskip("safetensors.torch")
    q = tmp_path / "model.safetensors"
    st.save_file({k: v.contiguous() for k, v in a.items()}, str(q))
    b = load_vision_checkpoint(str(q))
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
