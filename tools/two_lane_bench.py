"""Experiment: two independent library contexts ("lanes") on two streams, alternate batches between them with no
cross-stream synchronisation, so one lane's LayerNorm/attention/small kernels and GEMM tails overlap the other lane's
GEMMs.  Prints samples/s for 1 lane and 2 lanes.   python tools/two_lane_bench.py [S] [steps]"""
import ctypes as C
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"))
import torch  # noqa: E402
from ttl_b200 import Engine, Hparams  # noqa: E402
from ttl_b200 import _lib as L  # noqa: E402
from ttl_b200.synthetic import synthetic_vit_weights, synthetic_lora_init, synthetic_text_features  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 3
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
w = synthetic_vit_weights("ViT-B/16", seed=1234)
text = synthetic_text_features(1000, 512, seed=11)
lora = synthetic_lora_init("ViT-B/16", rank=16, layers=(9, 11), seed=0)


def mk():
    e = Engine("ViT-B/16", max_views=64, max_classes=1000, max_samples=S)
    e.load_weights(w)
    e.set_text_features(text, math.log(100.0))
    e.set_lora_init(lora)
    return e


lanes = [mk(), mk()]
hp = Hparams(head="tpt").to_c()
ring = [torch.randn(S, 64, 3, 224, 224, device="cuda") for _ in range(4)]
outs = [torch.empty(S, 1000, device="cuda") for _ in range(2)]
torch.cuda.synchronize()


def run(n_lanes, n):
    o = [L.TtlOutputs() for _ in range(2)]
    for k in range(2):
        o[k].pred_logits = outs[k].data_ptr()
    for i in range(n):
        e = lanes[i % n_lanes]
        L.check(e.lib.ttl_adapt_predict_batch(e.ctx, ring[i % 4].data_ptr(), S, 64, C.byref(hp), None, C.byref(o[i % n_lanes]),
                                              C.c_void_p(e.stream.cuda_stream)), e.ctx)


for n_lanes in (1, 2, 1, 2):
    run(n_lanes, 8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(n_lanes, steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"lanes={n_lanes} S={S}: {steps * S / dt:.1f} samples/s ({dt / steps * 1e3:.2f} ms/batch)", flush=True)
