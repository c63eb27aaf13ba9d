"""Numerical comparison of the attention forward kernels against an fp64 reference (development aid).
python tools/attn_cmp.py   (TTL_ATTN=mma|tc selects the kernel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_util as gu
lib = gu.lib()
torch.manual_seed(0)
for (V, tokens, heads, qs) in ((4, 197, 12, 1.5), (1, 197, 12, 1.0), (4, 197, 12, 3.0)):
    d = heads * 64
    qkv = (torch.randn(V * tokens, 3 * d, device="cuda") * qs)
    qkv[:, 2 * d:] += 0.5          # V with a mean, so that a mis-weighted softmax shows up as a bias
    qkv = qkv.bfloat16()
    out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(V, heads, tokens, device="cuda")
    gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
    torch.cuda.synchronize()
    q, k, v = qkv.double().view(V, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = torch.softmax(s, -1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(V * tokens, d)
    err = out.double() - ref
    rows = err.view(V, tokens, d)
    print(os.environ.get("TTL_ATTN", "tc"), V, tokens, qs, "rel", float(err.norm() / ref.norm()), "bias", float(err.mean()),
          "row0 rel", float(rows[:, 0].norm() / ref.view(V, tokens, d)[:, 0].norm()),
          "rows>=192 rel", float(rows[:, 192:].norm() / ref.view(V, tokens, d)[:, 192:].norm()),
          "rows 128..191 rel", float(rows[:, 128:192].norm() / ref.view(V, tokens, d)[:, 128:192].norm()),
          "lse err", float((lse.double() - torch.logsumexp(s, -1)).abs().max()))
    # weight on the last keys: does the tail block carry its share?
    w_tail = p[..., 192:].sum(-1).mean()
    print("   mean attention mass on keys >= 192:", float(w_tail))
