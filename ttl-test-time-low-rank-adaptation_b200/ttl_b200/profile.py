"""Live roofline measurement of the dominant kernel (the tcgen05 GEMM) for bench.py.

Every GEMM launch of a few adapted samples is bracketed by CUDA events on the launching stream (library side,
ttl_profile_gemm); algorithmic FLOPs per launch = 2*M*N*K with the true (unpadded) M.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from typing import Dict, List


from . import _lib as L

EPI_NAMES = {0: "bias->bf16", 1: "bias+QuickGELU->bf16", 2: "bias+residual->f32", 3: "patch-embed scatter->f32",
             4: "->f32", 5: "dQuickGELU->bf16"}


def read_records(eng) -> List[L.TtlGemmRecord]:
    cap = 1 << 16
    buf = (L.TtlGemmRecord * cap)()
    n = C.c_int32()
    L.check(eng.lib.ttl_profile_read(eng.ctx, buf, cap, C.byref(n)), eng.ctx)
    return [buf[i] for i in range(n.value)]


def gemm_roofline(eng, hp, ring, peaks: Dict[str, float], batches: int = 3, traffic=None) -> Dict:
    """`ring`: list of [S,V,3,size,size] device batches.  The dominant kernel is the CTA-pair tcgen05 GEMM of the
    S*V-view forward (every launch with M >= 4096 rows); the small-M launches of the 6-view backward and the 1-view
    prediction are reported in `all_gemm_launches` and `by_shape`."""
    S = int(ring[0].shape[0])
    L.check(eng.lib.ttl_profile_gemm(eng.ctx, 1), eng.ctx)
    try:
        eng.adapt_predict_batch(ring[0], hp, want=("pred_logits",))        # untimed warm-up in eager mode
        read_records(eng)
        for i in range(batches):
            eng.adapt_predict_batch(ring[(i + 1) % len(ring)], hp, want=("pred_logits",))
        recs = read_records(eng)
    finally:
        L.check(eng.lib.ttl_profile_gemm(eng.ctx, 0), eng.ctx)
    by = defaultdict(lambda: [0, 0.0, 0.0])
    flops = ms = big_flops = big_ms = 0.0
    n_big = 0
    for r in recs:
        f = 2.0 * r.M * r.N * r.K
        k = (r.M, r.N, r.K, r.epi)
        by[k][0] += 1
        by[k][1] += r.ms
        by[k][2] += f
        flops += f
        ms += r.ms
        if r.M >= 4096:
            big_flops += f
            big_ms += r.ms
            n_big += 1
    shapes = sorted(by.items(), key=lambda kv: -kv[1][1])[:6]
    achieved = big_flops / (big_ms * 1e-3) / 1e12 if big_ms > 0 else 0.0
    peak = peaks["tf_sus"]
    return {"bound": "tensor", "kernel": "gemm2_kernel<BLOCK_N,EPI> (cta_group::2 tcgen05 GEMM; all launches with M >= 4096 "
                                         "of the adapted batch)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic.get("bytes_per_launch_avg") if isinstance(traffic, dict) else None,
            "traffic_detail": traffic,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
            "launches_timed": n_big, "avg_launch_ms": big_ms / max(n_big, 1),
            "flop_per_launch_avg": big_flops / max(n_big, 1),
            "gemm_ms_per_sample": ms / (batches * S),
            "all_gemm_launches": {"launches": len(recs), "tflops": flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0,
                                  "ms_per_sample": ms / (batches * S)},
            "how": f"CUDA events around each GEMM launch on the launching stream, {batches} batches of {S} samples, eager "
                   f"pass right after the timed region; algorithmic FLOPs = 2*M*N*K (true M)",
            "by_shape": [{"M": k[0], "N": k[1], "K": k[2], "epilogue": EPI_NAMES.get(k[3], str(k[3])), "launches": v[0],
                          "avg_ms": v[1] / v[0], "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0}
                         for k, v in shapes]}
