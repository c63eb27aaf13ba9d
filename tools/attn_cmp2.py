import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_util as gu
lib = gu.lib()
torch.manual_seed(1)
heads, tokens = 12, 197
d = heads * 64
for V in (1, 2, 6, 64):
    for trial in range(2):
        base = torch.randn(V * tokens, 3 * d, device="cuda")
        # structured like a transformer layer: shared direction across tokens (mean offset) + token noise
        qkv = (base * 0.8 + torch.randn(1, 3 * d, device="cuda") * 1.5).bfloat16()
        out = torch.zeros(V * tokens, d, device="cuda", dtype=torch.bfloat16)
        gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), None, V, tokens, heads, 0.125, gu.stream()))
        torch.cuda.synchronize()
        q, k, v = qkv.double().view(V, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
        p = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1)
        ref = (p @ v).permute(0, 2, 1, 3).reshape(V * tokens, d)
        err = (out.double() - ref).view(V, tokens, heads, 64)
        refv = ref.view(V, tokens, heads, 64)
        per_head = ((err ** 2).sum(dim=(0, 1, 3)).sqrt() / (refv ** 2).sum(dim=(0, 1, 3)).sqrt()).cpu().numpy().round(4)
        print(os.environ.get("TTL_ATTN", "tc"), "V", V, "rel", float(err.norm() / refv.norm()), "rows<128", float(err[:, :128].norm() / refv[:, :128].norm()),
              "rows>=128", float(err[:, 128:].norm() / refv[:, 128:].norm()), "maxabs", float(err.abs().max()), "per-head", per_head)
