"""Where the energy of a step goes: each kernel class of the frozen pass run alone in a ~2.5 s loop at the bench's shapes (9 samples
x 64 views), board power and SM clock sampled through NVML every 10 ms.  Energy per launch = mean power x time per launch; with the
launches per step of the launch list this gives the energy share of every class next to its time share.
Usage: python tools/energy_probe.py [--seconds 2.5] > profiles/rN_energy_probe.json"""
import argparse, json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import pynvml
import gpu_util as gu
from ttl_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=2.5)
ap.add_argument("--samples", type=int, default=9)
args = ap.parse_args()
V, tokens, heads, d, F = 64 * args.samples, 197, 12, 768, 3072
M = V * tokens
lib = gu.lib()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.on, self.p, self.f = True, [], []

    def run(self):
        while self.on:
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            self.f.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.01)


def probe(name, fn, launches_per_step, flops=0.0, bytes_=0.0):
    for i in range(10):
        fn(i)
    torch.cuda.synchronize()
    # settle ~1 s, then measure
    t_end = time.perf_counter() + 1.0
    i = 0
    while time.perf_counter() < t_end:
        for _ in range(20):
            fn(i); i += 1
        torch.cuda.synchronize()
    s = Sampler(); s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    t_end = time.perf_counter() + args.seconds
    while time.perf_counter() < t_end:
        for _ in range(20):
            fn(i); i += 1; n += 1
        torch.cuda.synchronize()      # bounds the queue depth; the gap is ~20 us per 20 launches
    e1.record(); torch.cuda.synchronize()
    s.on = False; s.join()
    us = e0.elapsed_time(e1) * 1e3 / n
    p = sorted(s.p)[len(s.p) // 2]
    f = sorted(s.f)[len(s.f) // 2]
    rec = {"kernel": name, "us_per_launch": us, "power_w_median": p, "sm_mhz_median": f, "mj_per_launch": p * us * 1e-3,
           "launches_per_step": launches_per_step, "mj_per_step": p * us * 1e-3 * launches_per_step,
           "ms_per_step": us * 1e-3 * launches_per_step}
    if flops: rec["tflops"] = flops / us / 1e6
    if bytes_: rec["gbs"] = bytes_ / us / 1e3
    print(json.dumps(rec), flush=True)
    return rec


ring = 3
recs = []
# idle power
time.sleep(1.0)
idle = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
# GEMMs of one frozen layer
for name, N, K, epi in (("qkv", 2304, 768, L.EPI_BF16), ("out-proj", 768, 768, L.EPI_RESID_F32), ("fc1", 3072, 768, L.EPI_GELU),
                        ("fc2", 768, 3072, L.EPI_RESID_F32)):
    As = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(ring)]
    B = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    f32 = epi == L.EPI_RESID_F32
    outs = [torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16) for _ in range(ring)]
    res = [torch.randn(M, N, device="cuda") for _ in range(ring)] if f32 else [None] * ring

    def run(i, As=As, B=B, bias=bias, outs=outs, res=res, N=N, K=K, epi=epi):
        gu.ok(lib.ttl_op_gemm(gu.ptr(As[i % ring]), gu.ptr(B), None, None, M, N, K, 0, epi, gu.ptr(bias), gu.ptr(outs[i % ring]),
                              None, gu.ptr(res[i % ring]), None, None, 0, 0, gu.stream()))
    recs.append(probe("gemm " + name, run, 12, flops=2.0 * M * N * K))
    del As, outs, res
    torch.cuda.empty_cache()
# LayerNorm
xs = [torch.randn(M, d, device="cuda") for _ in range(ring)]
ys = [torch.empty(M, d, device="cuda", dtype=torch.bfloat16) for _ in range(ring)]
gam, bet = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
recs.append(probe("layernorm", lambda i: gu.ok(lib.ttl_op_layernorm(gu.ptr(xs[i % ring]), gu.ptr(ys[i % ring]), gu.ptr(gam), gu.ptr(bet), M, d, 1e-5, gu.stream())),
                  24, bytes_=M * d * 6.0))
del xs, ys
# attention forward
qkvs = [(torch.randn(M, 3 * d, device="cuda") * 1.5).bfloat16() for _ in range(ring)]
aos = [torch.empty(M, d, device="cuda", dtype=torch.bfloat16) for _ in range(ring)]
recs.append(probe("attention fwd", lambda i: gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkvs[i % ring]), gu.ptr(aos[i % ring]), None, V, tokens, heads, 0.125, gu.stream())),
                  12, flops=4.0 * tokens * tokens * 64 * heads * V, bytes_=M * d * 8.0))
tot_ms = sum(r["ms_per_step"] for r in recs)
tot_mj = sum(r["mj_per_step"] for r in recs)
summary = {"idle_power_w": idle, "frozen_pass_ms_per_step_sum": tot_ms, "frozen_pass_j_per_step_sum": tot_mj * 1e-3,
           "shares": {r["kernel"]: {"time": r["ms_per_step"] / tot_ms, "energy": r["mj_per_step"] / tot_mj} for r in recs},
           "note": "each class alone in a loop at M = %d rows; a step = %d samples; 12 frozen-pass layers (layer 11 runs CLS-only in the step)" % (M, args.samples)}
print(json.dumps(summary), flush=True)
