"""Golden token ids from the reference's own tokenizer (clip/simple_tokenizer.py + clip.tokenize, clip/clip.py:196-232) for
a list of prompts -> tests/golden/tokenizer_golden.json.  Dev-container only.  Usage: python oracle/make_golden_tokens.py"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim as R  # noqa: E402

PROMPTS = ["a photo of a airplane.", "a photo of a great white shark.", "a photo of a Tench, Tinca tinca.",
           "a photo of a jack-o'-lantern.", "a photo of a three-toed sloth.", "A   photo\tof a  CD player .",
           "a photo of a crème brûlée.", "it's the dog's toy, we've 2 of them: 42!", "a photo of a hen-of-the-woods.",
           "a photo of a &amp; symbol", "itap of a 12-year-old's birthday cake", "a photo of a toilet paper.",
           "a photo of a T-shirt.", "a photo of a Übergröße straße", "日本語 のテキスト", "", "a_photo_of_a street_sign."]


def main() -> None:
    R.install(None)
    import clip
    toks = clip.tokenize(PROMPTS)
    rec = {"prompts": PROMPTS, "tokens": toks.tolist()}
    path = os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json")
    with open(path, "w") as f:
        json.dump(rec, f, ensure_ascii=False)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
