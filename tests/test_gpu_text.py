"""Text tower on the device (csrc/text.cu) against the fp32 CPU oracle (oracle/text_oracle.py, pinned
to HF transformers in tests/test_text_oracle.py).  bf16 GEMM operands, fp32 statistics: class features (unit vectors)
within 1e-2 relative, the north-star tolerance for bf16."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import text_oracle as TO  # noqa: E402


def _run(arch_name, geom_key, n, max_prompts):
    from ttl_b200.text import TextEncoder
    a = TO.TEXT_ARCHS[arch_name]
    w = TO.make_synthetic_text_weights(a, seed=5)
    tokens = TO.make_synthetic_tokens(n, a, seed=3)
    ref = TO.text_forward(a, w, tokens)
    enc = TextEncoder(geom_key, max_prompts=max_prompts)
    try:
        enc.load_weights(w)
        got = enc.encode(tokens)
        again = enc.encode(tokens[: max(1, n // 2)])
    finally:
        enc.close()
    assert got.shape == ref.shape
    rel = float((got - ref).norm() / ref.norm())
    cos = float((got * ref).sum(dim=1).min())
    k = max(1, n // 2)      # another chunking picks other GEMM tile shapes / epilogue variants: same features within bf16 noise
    assert float((again - got[:k]).norm() / got[:k].norm()) < 5e-3
    return rel, cos


def test_text_tower_tiny_vs_oracle():
    rel, cos = _run("tiny", "tiny", 13, 5)       # 13 prompts in chunks of 5: ragged last chunk
    assert rel < 1e-2 and cos > 0.9999, (rel, cos)


def test_text_tower_b16_vs_oracle():
    rel, cos = _run("ViT-B/16", "ViT-B/16", 24, 16)
    assert rel < 1e-2 and cos > 0.9999, (rel, cos)


def test_text_tower_rejects_bad_shapes():
    from ttl_b200.text import TextEncoder
    enc = TextEncoder("tiny", max_prompts=4)
    try:
        with pytest.raises(ValueError):
            enc.encode(np.zeros((2, 5), dtype=np.int32))
        with pytest.raises(RuntimeError):
            enc._set(0, 999, np.zeros(4, dtype=np.float32))
    finally:
        enc.close()


def test_module_builds_class_features_with_the_text_tower():
    """ClipTestTimeTuning with text-tower weights: get_text_features() = tokenizer -> device text tower (once per class-name
    set), equal to the oracle's features within bf16 tolerance; reset_classnames recomputes them."""
    import zlib
    from clip.custom_clip import ClipTestTimeTuning
    from ttl_b200.synthetic import synthetic_vit_weights
    a = TO.TEXT_ARCHS["ViT-B/16"]
    tw = TO.make_synthetic_text_weights(a, seed=5)

    def fake_tokenizer(prompts):        # stands in for the BPE table (data, not shipped): stable ids from the characters
        out = torch.zeros(len(prompts), a.context, dtype=torch.long)
        for i, p in enumerate(prompts):
            ids = [1 + zlib.crc32(w.encode()) % (a.vocab - 3) for w in p.replace(".", " .").split()]
            row = [a.vocab - 2] + ids + [a.vocab - 1]
            out[i, :len(row)] = torch.tensor(row)
        return out

    names = ["airplane", "automobile", "bird", "cat", "deer"]
    m = ClipTestTimeTuning(0, names, None, arch="ViT-B/16", layer_range=[9, 11], lora_encoder="image", max_views=4,
                           weights=synthetic_vit_weights("ViT-B/16", seed=1234), text_weights=tw, tokenizer=fake_tokenizer)
    try:
        prompts = [f"a photo of a {n}." for n in names]
        assert m.prompt_learner.prompts == prompts
        ref = TO.text_forward(a, tw, fake_tokenizer(prompts))
        got = m.get_text_features().cpu()
        assert got.shape == (5, 512) and float((got - ref).norm() / ref.norm()) < 1e-2
        assert torch.equal(m.prompt_learner.tokenized_prompts.cpu(), fake_tokenizer(prompts))
        m.reset_classnames(["dog", "frog", "horse"], "ViT-B/16")
        ref2 = TO.text_forward(a, tw, fake_tokenizer([f"a photo of a {n}." for n in ["dog", "frog", "horse"]]))
        got2 = m.get_text_features().cpu()
        assert got2.shape == (3, 512) and float((got2 - ref2).norm() / ref2.norm()) < 1e-2
        logits = m(torch.randn(2, 3, 224, 224, device="cuda"))
        assert logits.shape == (2, 3)
    finally:
        m.engine.close()
