import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_util as gu
lib = gu.lib()
V, tokens, heads = int(os.environ.get("BWD_V", "576")), 197, 12
d = heads * 64
qkv = (torch.randn(V * tokens, 3 * d, device="cuda") * 1.5).bfloat16()
out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(V, heads, tokens, device="cuda")
dout = torch.randn(V * tokens, d, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
for i in range(5):
    gu.ok(lib.ttl_op_attention_bwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(dout), gu.ptr(lse), gu.ptr(dqkv), V, tokens, heads, 0.125, gu.stream()))
torch.cuda.synchronize()
