"""Experiment: one GEMM as TWO concurrent launches -- the 4-CTA-cluster kernel (multicast B tiles, 33 co-resident clusters =
132 SMs) on the first rows and the CTA-pair kernel, capped at the 8 pairs the cluster placement strands, on the rest --
against the pair kernel alone.  Sustained timing (power-capped), fork/join with events per iteration.
Usage: python tools/gemm_split_bench.py [--iters 200] [--M 75648] [--plain]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_util as gu
from ttl_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--M", type=int, default=75648)
ap.add_argument("--only", default="")
ap.add_argument("--plain", action="store_true")
ap.add_argument("--pairs", type=int, default=8, help="pairs left to the CTA-pair kernel")
ap.add_argument("--quads", type=int, default=33)
args = ap.parse_args()
M = args.M
shapes = [("qkv", 2304, 768, L.EPI_BF16), ("out", 768, 768, L.EPI_RESID_F32), ("fc1", 3072, 768, L.EPI_GELU),
          ("fc2", 768, 3072, L.EPI_RESID_F32)]
if args.plain:
    shapes = [(n, N, K, L.EPI_BF16) for n, N, K, _ in shapes]
lib = gu.lib()
s_main = torch.cuda.current_stream()
s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()


def sptr(s):
    import ctypes
    return ctypes.c_void_p(s.cuda_stream)


for name, N, K, epi in shapes:
    if args.only and name not in args.only.split(","):
        continue
    ring = 4
    As = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(ring)]
    B = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    f32 = epi in (L.EPI_RESID_F32, L.EPI_F32)
    outs = [torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16) for _ in range(ring)]
    res = [torch.randn(M, N, device="cuda") for _ in range(ring)] if epi == L.EPI_RESID_F32 else [None] * ring
    n_tiles = N // 256
    blocks = (M + 255) // 256

    def gemm(i, r0, r1, code, stream):
        a, o = As[i % ring][r0:r1], outs[i % ring][r0:r1]
        r = res[i % ring][r0:r1] if res[i % ring] is not None else None
        gu.ok(lib.ttl_op_gemm(gu.ptr(a), gu.ptr(B), None, None, r1 - r0, N, K, 0, epi, gu.ptr(bias), gu.ptr(o), None,
                              gu.ptr(r) if r is not None else None, None, None, 0, code, stream))

    def timeit(fn):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / args.iters

    def check(i):
        ref = As[i % ring].float() @ B.float().t() + bias
        if epi == L.EPI_RESID_F32:
            ref = ref + res[i % ring]
        if epi == L.EPI_GELU:
            ref = ref * torch.sigmoid(1.702 * ref)
        return gu.rel_err(outs[i % ring], ref)

    us = timeit(lambda i: gemm(i, 0, M, 1256, gu.stream()))
    print(f"{name:4s} pair kernel alone            {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  err {check(args.iters - 1):.1e}", flush=True)
    us = timeit(lambda i: gemm(i, 0, M, 2256, gu.stream()))
    print(f"{name:4s} 4-CTA clusters alone         {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  err {check(args.iters - 1):.1e}", flush=True)
    # candidate splits: x row blocks of 512 to the clusters, the rest to the pairs; model cost = max(rounds4 / 1.08, rounds2)
    cands = []
    for x in range(blocks // 2 - 40, blocks // 2 + 1):
        rest = blocks - 2 * x
        if rest <= 0:
            continue
        r4 = -(-x * n_tiles // args.quads)
        r2 = -(-rest * n_tiles // args.pairs)
        cands.append((max(r4 / 1.08, r2), x))
    cands.sort()
    for cost, x in cands[:3]:
        m4 = min(M, x * 512)

        def split(i, m4=m4):
            ev = torch.cuda.Event()
            ev.record(s_main)
            s_a.wait_event(ev)
            s_b.wait_event(ev)
            gemm(i, 0, m4, 2256, sptr(s_a))
            gemm(i, m4, M, 1256 + 100000 * args.pairs, sptr(s_b))
            ea, eb = torch.cuda.Event(), torch.cuda.Event()
            ea.record(s_a)
            eb.record(s_b)
            s_main.wait_event(ea)
            s_main.wait_event(eb)

        us = timeit(split)
        print(f"{name:4s} split rows {m4:6d} | {M - m4:5d} (model {cost:5.1f} rounds)  {us:8.1f} us  "
              f"{2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  err {check(args.iters - 1):.1e}", flush=True)
