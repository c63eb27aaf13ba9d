"""Micro-benchmark of the LoRA weight-gradient reduction (launch_skinny_reduce) through the C ABI: dB = s dY^T (X A^T) over all rows.
Usage: python tools/skinny_bench.py   (TTL_SKINNY=mma for the mma.sync kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_util as gu

lib = gu.lib()
for M in (10638, 113472):
    wide = (torch.randn(M, 2304, device="cuda")).bfloat16()
    narrow = (torch.randn(M, 64, device="cuda")).bfloat16()
    ws = torch.empty((M + 127) // 128 * 768 * 32 + 768 * 32, device="cuda")
    out = torch.empty(768, 16, device="cuda")
    run = lambda: gu.ok(lib.ttl_op_skinny_reduce(gu.ptr(wide), 2304, 768, gu.ptr(narrow), 64, 16, M, 2.0, gu.ptr(out), 0, gu.ptr(ws), gu.stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ref = 2.0 * wide[:, :768].float().t() @ narrow[:, :16].float()
    err = float((out - ref).norm() / ref.norm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"skinny reduce M={M}: {us:7.1f} us, {M * 768 * 2 / us / 1e3:7.1f} GB/s of dY, rel err {err:.1e}", flush=True)
