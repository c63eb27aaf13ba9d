"""Byte-level BPE tokenizer of CLIP (host side of the class-feature builder).

Produces what the reference's `clip.tokenize` (clip/clip.py:196-232 over clip/simple_tokenizer.py:62-132) produces:
int tensors [n, 77] = <|startoftext|>, BPE ids of the cleaned lower-cased text, <|endoftext|>, zero padding.  The merge
table is data, not code: point `bpe_path` (or $TTL_BPE_PATH) at the `bpe_simple_vocab_16e6.txt.gz` that ships with
CLIP / with the reference (clip/bpe_simple_vocab_16e6.txt.gz); it is not vendored here.
`ftfy.fix_text` is applied when ftfy is installed (the reference requires it); plain ASCII prompts such as the class
templates are unaffected by it."""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import regex
import torch

SOT, EOT = "<|startoftext|>", "<|endoftext|>"
N_MERGES = 49152 - 256 - 2
_SPLIT = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                       regex.IGNORECASE)


def default_bpe_path() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.environ.get("TTL_BPE_PATH"), os.path.join(here, "bpe_simple_vocab_16e6.txt.gz"),
             os.path.join(os.path.dirname(here), "clip", "bpe_simple_vocab_16e6.txt.gz")]
    for c in cands:
        if c and os.path.exists(c):
            return c
    raise FileNotFoundError("CLIP BPE merge table not found: set TTL_BPE_PATH to bpe_simple_vocab_16e6.txt.gz")


@lru_cache()
def byte_alphabet() -> Dict[int, str]:
    """Reversible byte -> printable-character table of GPT-2 style BPE: the 188 printable latin-1 bytes stand for
    themselves, the other 68 are moved to code points 256.."""
    keep = set(range(ord("!"), ord("~") + 1)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, shift = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + shift)
            shift += 1
    return table


def _clean(text: str) -> str:
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(html.unescape(text)).strip()
    return regex.sub(r"\s+", " ", text).strip()


class SimpleTokenizer:
    def __init__(self, bpe_path: str | None = None):
        path = bpe_path or default_bpe_path()
        # the OpenAI release gzips the table; HF checkpoints carry the same lines as plain `merges.txt`
        with (gzip.open(path, "rt", encoding="utf-8") if path.endswith(".gz") else open(path, "rt", encoding="utf-8")) as f:
            lines = f.read().split("\n")
        merges: List[Tuple[str, str]] = [tuple(l.split()) for l in lines[1:1 + N_MERGES] if l.strip()]   # line 0 is a header
        alphabet = list(byte_alphabet().values())
        # GPT-2 orders its alphabet "printable bytes first (in byte order), then the shifted ones"
        printable = [c for c in alphabet if ord(c) < 256]
        shifted = [c for c in alphabet if ord(c) >= 256]
        symbols = printable + shifted
        vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.b2u = byte_alphabet()
        self.u2b = {u: b for b, u in self.b2u.items()}
        self._cache: Dict[str, List[str]] = {SOT: [SOT], EOT: [EOT]}

    # ---- BPE of one pre-token
    def _bpe(self, token: str) -> List[str]:
        hit = self._cache.get(token)
        if hit is not None:
            return hit
        parts = list(token[:-1]) + [token[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = -1, None
            for i in range(len(parts) - 1):
                r = self.rank.get((parts[i], parts[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = parts[best], parts[best + 1]
            merged, i = [], 0
            while i < len(parts):          # merge EVERY occurrence of the winning pair, left to right
                if i + 1 < len(parts) and parts[i] == a and parts[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        self._cache[token] = parts
        return parts

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for tok in _SPLIT.findall(_clean(text).lower()):
            mapped = "".join(self.b2u[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self._bpe(mapped))
        return ids

    def decode(self, ids: Iterable[int]) -> str:
        raw = bytearray()
        for i in ids:
            tok = self.decoder[int(i)]
            if tok in (SOT, EOT):
                raw.extend(tok.encode() + b" ")
                continue
            word_end = tok.endswith("</w>")
            raw.extend(self.u2b[c] for c in (tok[:-4] if word_end else tok))
            if word_end:
                raw.append(0x20)
        return raw.decode("utf-8", errors="replace")

    def __call__(self, texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
        """clip.tokenize: [n, context_length] int64."""
        if isinstance(texts, str):
            texts = [texts]
        sot, eot = self.encoder[SOT], self.encoder[EOT]
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [sot] + self.encode(t) + [eot]
            if len(ids) > context_length:
                if not truncate:
                    raise RuntimeError(f"Input {t} is too long for context length {context_length}")
                ids = ids[:context_length]
                ids[-1] = eot
            out[i, :len(ids)] = torch.tensor(ids)
        return out


def class_prompts(classnames: Sequence[str], ctx_init: str = "a_photo_of_a") -> List[str]:
    """The hand-crafted prompts of the TTL path (PromptLearner with ctx_init, clip/custom_clip.py:343-372)."""
    lead = ctx_init.replace("_", " ")
    return [f"{lead} {name.replace('_', ' ')}." for name in classnames]
