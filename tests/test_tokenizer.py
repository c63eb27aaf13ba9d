"""ttl_b200.tokenizer against golden token ids produced by the reference's own tokenizer (clip.tokenize over
clip/simple_tokenizer.py; oracle/make_golden_tokens.py).  Needs CLIP's BPE merge table, which is data and not vendored:
the test runs wherever $TTL_BPE_PATH or /root/reference/clip/bpe_simple_vocab_16e6.txt.gz exists and skips elsewhere."""
import json
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tokenizer_golden.json")
BPE = os.environ.get("TTL_BPE_PATH") or "/root/reference/clip/bpe_simple_vocab_16e6.txt.gz"


def test_byte_alphabet_is_a_bijection_onto_printables():
    from ttl_b200.tokenizer import byte_alphabet
    t = byte_alphabet()
    assert len(t) == 256 and len(set(t.values())) == 256
    assert t[ord("a")] == "a" and t[ord(" ")] == chr(256 + 32) and all(not c.isspace() for c in t.values())


def test_class_prompts_follow_the_template():
    from ttl_b200.tokenizer import class_prompts
    assert class_prompts(["airplane", "street_sign"]) == ["a photo of a airplane.", "a photo of a street sign."]


@pytest.mark.skipif(not os.path.exists(BPE), reason="CLIP BPE merge table not available")
def test_tokens_match_the_reference_tokenizer():
    from ttl_b200.tokenizer import SimpleTokenizer
    g = json.load(open(GOLD))
    tk = SimpleTokenizer(BPE)
    got = tk(g["prompts"])
    assert got.dtype == torch.long and got.shape == (len(g["prompts"]), 77)
    assert got.tolist() == g["tokens"]
    assert tk.decode(got[0][1:7].tolist()).strip() == "a photo of a airplane ."
    with pytest.raises(RuntimeError):
        tk("word " * 100)
    long = tk("word " * 100, truncate=True)
    assert int(long[0, -1]) == tk.encoder["<|endoftext|>"] and int(long[0].argmax()) == 76
