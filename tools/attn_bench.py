"""Micro-benchmark of the fused attention kernels through the C ABI.  Usage: python tools/attn_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_util as gu

lib = gu.lib()
FWD = ((64, 197, 12), (192, 197, 12), (6, 197, 12), (1, 197, 12), (64, 257, 16))
if os.environ.get("ATTN_BENCH_FWD"):      # e.g. ATTN_BENCH_FWD=192,576: only these view counts (197 tokens x 12 heads unless ATTN_BENCH_TOKENS / _HEADS say otherwise), no backward
    _tk, _hd = int(os.environ.get("ATTN_BENCH_TOKENS", "197")), int(os.environ.get("ATTN_BENCH_HEADS", "12"))      # 257 / 16: ViT-L/14
    FWD = tuple((int(v), _tk, _hd) for v in os.environ["ATTN_BENCH_FWD"].split(","))
for (V, tokens, heads) in FWD:
    d = heads * 64
    ring = 4
    qkvs = [(torch.randn(V * tokens, 3 * d, device="cuda") * 1.5).bfloat16() for _ in range(ring)]
    outs = [torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16) for _ in range(ring)]
    def run(i):
        gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkvs[i % ring]), gu.ptr(outs[i % ring]), None, V, tokens, heads, 0.125, gu.stream()))
    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 40
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 4.0 * tokens * tokens * 64 * heads * V
    by = V * tokens * d * 2 * 4
    print(f"attention fwd V={V} tokens={tokens} heads={heads}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s", flush=True)

for (V, tokens, heads) in (() if os.environ.get("ATTN_BENCH_FWD") else ((18, 197, 12), (6, 197, 12), (54, 197, 12), (64, 197, 12), (576, 197, 12))):
    d = heads * 64
    qkv = (torch.randn(V * tokens, 3 * d, device="cuda") * 1.5).bfloat16()
    out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(V, heads, tokens, device="cuda")
    dout = torch.randn(V * tokens, d, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
    def runb():
        gu.ok(lib.ttl_op_attention_bwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(dout), gu.ptr(lse), gu.ptr(dqkv), V, tokens, heads, 0.125, gu.stream()))
    for i in range(3):
        runb()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        runb()
    e1.record()
    torch.cuda.synchronize()
    print(f"attention bwd V={V}: {e0.elapsed_time(e1) * 1e3 / 20:8.1f} us", flush=True)
