#!/usr/bin/env python
"""What the same B200 gives WITHOUT this library: the TTL per-sample loop (ttl.py:338-352, north-star head) written in
stock PyTorch -- nn.Linear / F.scaled_dot_product_attention / autograd / torch.optim.AdamW under torch.autocast(bf16) --
with the class features cached (the reference re-runs its text tower twice per sample on top of this, so this line is
FASTER than the real reference would be on the GPU).  SURVEY.md 8d "stock PyTorch bf16 path on the same B200".
Timing tool only: random-init weights, synthetic views, no parity claim; imports nothing from oracle/ or the library.

    python tools/torch_gpu_baseline.py [--samples 24] [--warmup 4] [--classes 1000] [--views 64]
Prints one JSON line.
"""
from __future__ import annotations

import argparse
import copy
import json
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class LoraLinear(nn.Module):
    def __init__(self, d, r=16, alpha=32.0):
        super().__init__()
        self.base = nn.Linear(d, d)
        self.A = nn.Linear(d, r, bias=False)
        self.B = nn.Linear(r, d, bias=False)
        self.s = alpha / r
        nn.init.xavier_normal_(self.A.weight)
        nn.init.zeros_(self.B.weight)

    def forward(self, x):
        return self.base(x) + self.B(self.A(x)) * self.s


class Layer(nn.Module):
    def __init__(self, d, heads, lora):
        super().__init__()
        self.h = heads
        self.ln1, self.ln2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.q = LoraLinear(d) if lora else nn.Linear(d, d)
        self.k = nn.Linear(d, d)
        self.v = LoraLinear(d) if lora else nn.Linear(d, d)
        self.o = nn.Linear(d, d)
        self.fc1, self.fc2 = nn.Linear(d, 4 * d), nn.Linear(4 * d, d)

    def forward(self, x):
        b, n, d = x.shape
        h = self.ln1(x)
        sp = lambda t: t.view(b, n, self.h, d // self.h).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(self.q(h)), sp(self.k(h)), sp(self.v(h)))
        x = x + self.o(a.transpose(1, 2).reshape(b, n, d))
        z = self.fc1(self.ln2(x))
        return x + self.fc2(z * torch.sigmoid(1.702 * z))


class Vit(nn.Module):
    def __init__(self, d=768, layers=12, heads=12, patch=16, size=224, proj=512, lora_from=9):
        super().__init__()
        self.conv = nn.Conv2d(3, d, patch, patch, bias=False)
        self.cls = nn.Parameter(torch.randn(d) * d ** -0.5)
        self.pos = nn.Parameter(torch.randn((size // patch) ** 2 + 1, d) * 0.02)
        self.pre, self.post = nn.LayerNorm(d), nn.LayerNorm(d)
        self.layers = nn.ModuleList([Layer(d, heads, i >= lora_from) for i in range(layers)])
        self.proj = nn.Linear(d, proj, bias=False)

    def forward(self, img):
        x = self.conv(img).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls.expand(x.shape[0], 1, -1).to(x.dtype), x], 1) + self.pos
        x = self.pre(x)
        for l in self.layers:
            x = l(x)
        return self.proj(self.post(x[:, 0]))


def run(samples: int = 24, warmup: int = 4, classes: int = 1000, views: int = 64, geo: dict | None = None) -> dict:
    """Time `samples` adapted samples of the stock-PyTorch loop on the current CUDA device; returns the JSON-able record."""
    a = argparse.Namespace(samples=samples, warmup=warmup, classes=classes, views=views)
    rng_state = torch.random.get_rng_state()
    torch.manual_seed(0)
    dev = "cuda"
    m = Vit(**(geo or {})).to(dev).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    lora = [p for l in m.layers for mod in (l.q, l.v) if isinstance(mod, LoraLinear) for p in (mod.A.weight, mod.B.weight)]
    for p in lora:
        p.requires_grad_(True)
    init = [p.detach().clone() for p in lora]
    text = F.normalize(torch.randn(a.classes, m.proj.out_features, device=dev), dim=-1)
    opt = torch.optim.AdamW(lora, lr=5e-3)
    opt_state = copy.deepcopy(opt.state_dict())
    ring = [torch.randn(a.views, 3, 224, 224, device=dev) for _ in range(8)]      # 8 x 38.5 MB > L2

    def logits_of(img):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f = m(img)
        f = f.float()
        return 100.0 * F.normalize(f, dim=-1) @ text.t()

    def one(img):
        with torch.no_grad():
            for p, p0 in zip(lora, init):
                p.copy_(p0)
        opt.load_state_dict(opt_state)
        out = logits_of(img)
        ent = -(out.softmax(1) * out.log_softmax(1)).sum(1)
        idx = torch.argsort(ent)[: int(a.views * 0.1)]
        lp = out[idx].log_softmax(1)
        avg = torch.logsumexp(lp, 0) - math.log(lp.shape[0])
        loss = -(avg * avg.exp()).sum()
        opt.zero_grad()
        loss.backward()
        opt.step()
        with torch.no_grad():
            return logits_of(img[:1]).argmax()

    for i in range(a.warmup):
        one(ring[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.samples):
        one(ring[i % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    torch.random.set_rng_state(rng_state)
    del m, opt, ring
    torch.cuda.empty_cache()
    return {"what": "stock PyTorch bf16-autocast TTL loop on this GPU (nn.Linear / SDPA / autograd through all 64 views / "
                    "torch.optim.AdamW; class features cached, so faster than the real reference would be)",
            "value": a.samples / (ms * 1e-3), "unit": "samples/s", "ms_per_sample": ms / a.samples, "samples": a.samples,
            "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--views", type=int, default=64)
    a = ap.parse_args()
    print(json.dumps(run(a.samples, a.warmup, a.classes, a.views)))


if __name__ == "__main__":
    main()
