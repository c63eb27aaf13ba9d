"""Live roofline measurement of the dominant kernel (the tcgen05 GEMM) for bench.py.

Every GEMM launch of a few adapted samples is bracketed by CUDA events on the launching stream (library side,
ttl_profile_gemm); algorithmic FLOPs per launch = 2*M*N*K with the true (unpadded) M.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from typing import Dict, List

import torch

from . import _lib as L

EPI_NAMES = {0: "bias->bf16", 1: "bias+QuickGELU->bf16", 2: "bias+residual->f32", 3: "patch-embed scatter->f32",
             4: "->f32", 5: "dQuickGELU->bf16"}


def read_records(eng) -> List[L.TtlGemmRecord]:
    cap = 1 << 16
    buf = (L.TtlGemmRecord * cap)()
    n = C.c_int32()
    L.check(eng.lib.ttl_profile_read(eng.ctx, buf, cap, C.byref(n)), eng.ctx)
    return [buf[i] for i in range(n.value)]


def gemm_roofline(eng, hp, ring, peaks: Dict[str, float], samples: int = 4) -> Dict:
    L.check(eng.lib.ttl_profile_gemm(eng.ctx, 1), eng.ctx)
    try:
        eng.adapt_predict(ring[0], hp, want=("pred_logits",))        # untimed warm-up in eager mode
        read_records(eng)
        for i in range(samples):
            eng.adapt_predict(ring[(i + 1) % len(ring)], hp, want=("pred_logits",))
        recs = read_records(eng)
    finally:
        L.check(eng.lib.ttl_profile_gemm(eng.ctx, 0), eng.ctx)
    by = defaultdict(lambda: [0, 0.0, 0.0])
    flops = ms = 0.0
    for r in recs:
        f = 2.0 * r.M * r.N * r.K
        k = (r.M, r.N, r.K, r.epi)
        by[k][0] += 1
        by[k][1] += r.ms
        by[k][2] += f
        flops += f
        ms += r.ms
    shapes = sorted(by.items(), key=lambda kv: -kv[1][1])[:6]
    achieved = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    peak = peaks["tf_sus"]
    return {"bound": "tensor", "kernel": "gemm_tcgen05_kernel<BLOCK_N,EPI> (all launches of the adapted sample)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
            "launches_timed": len(recs), "avg_launch_ms": ms / max(len(recs), 1),
            "gemm_ms_per_sample": ms / samples,
            "how": f"CUDA events around each GEMM launch on the launching stream, {samples} samples, eager pass right "
                   f"after the timed region; algorithmic FLOPs = 2*M*N*K",
            "by_shape": [{"M": k[0], "N": k[1], "K": k[2], "epilogue": EPI_NAMES.get(k[3], str(k[3])), "launches": v[0],
                          "avg_ms": v[1] / v[0], "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0}
                         for k, v in shapes]}
