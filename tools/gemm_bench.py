"""Micro-benchmark of the tcgen05 GEMM kernels through the C ABI (ttl_op_gemm): CUDA-event timing per shape/tile
configuration, inputs rotated through a ring larger than L2.  Usage: python tools/gemm_bench.py [--iters 50]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_util as gu
from ttl_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--M", type=int, default=12608)
ap.add_argument("--only", default="")
ap.add_argument("--plain", action="store_true", help="bias -> bf16 epilogue for every shape")
ap.add_argument("--cublas", action="store_true", help="also time torch.addmm (cuBLASLt) at the same shapes")
args = ap.parse_args()
M = args.M
shapes = [("qkv", 2304, 768, L.EPI_BF16), ("out", 768, 768, L.EPI_RESID_F32), ("fc1", 3072, 768, L.EPI_GELU),
          ("fc2", 768, 3072, L.EPI_RESID_F32)]
if args.plain:     # every shape with the plain bias -> bf16 epilogue: isolates the main loop from the epilogue's traffic
    shapes = [(n, N, K, L.EPI_BF16) for n, N, K, _ in shapes]
cfgs = [int(c) for c in os.environ.get("CFGS", "0,256,1256,1192,1128").split(",")]
lib = gu.lib()
for name, N, K, epi in shapes:
    if args.only and name not in args.only.split(","):
        continue
    ring = 6
    As = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(ring)]
    B = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    f32 = epi in (L.EPI_RESID_F32, L.EPI_F32)
    outs = [torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16) for _ in range(ring)]
    res = [torch.randn(M, N, device="cuda") for _ in range(ring)] if epi == L.EPI_RESID_F32 else [None] * ring
    if args.cublas:   # library yardstick at the SAME shape: cuBLASLt bf16 GEMM + bias -> bf16 (no GELU / residual / fp32 output)
        Bb = bias.bfloat16()
        o = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(ring)]
        for i in range(5):
            torch.addmm(Bb, As[i % ring], B.t(), out=o[i % ring])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            torch.addmm(Bb, As[i % ring], B.t(), out=o[i % ring])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        print(f"{name:4s} M={M} N={N:5d} K={K:5d} cuBLAS addmm bf16->bf16  {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s", flush=True)
        del o
    for bn in cfgs:
        if bn >= 1000 and N % (bn % 1000):
            continue
        def run(i):
            gu.ok(lib.ttl_op_gemm(gu.ptr(As[i % ring]), gu.ptr(B), None, None, M, N, K, 0, epi, gu.ptr(bias), gu.ptr(outs[i % ring]),
                                  None, gu.ptr(res[i % ring]), None, None, 0, bn, gu.stream()))
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        ref = As[(args.iters - 1) % ring].float() @ B.float().t() + bias
        if epi == L.EPI_RESID_F32:
            ref = ref + res[(args.iters - 1) % ring]
        if epi == L.EPI_GELU:
            ref = ref * torch.sigmoid(1.702 * ref)
        err = gu.rel_err(outs[(args.iters - 1) % ring], ref)
        print(f"{name:4s} M={M} N={N:5d} K={K:5d} cfg={bn:5d}  {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  rel_err {err:.2e}", flush=True)
