# compute-sanitizer over the end-to-end driver (memcheck, synccheck) and over the attention kernels (racecheck).
export PYTHONPATH=.
O=gpurun_out/s43; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck python tests/sanitizer_driver_b16.py > $O/memcheck.log 2>&1; tail -3 $O/memcheck.log
timeout 900 compute-sanitizer --tool synccheck python tests/sanitizer_driver_b16.py > $O/synccheck.log 2>&1; tail -3 $O/synccheck.log
cat > /tmp/attn_small.py <<'PY'
import os, sys
sys.path[:0] = [".", "ttl-test-time-low-rank-adaptation_b200", "tests"]
import torch, gpu_util as gu
lib = gu.lib()
for (V, tokens, heads) in ((13, 197, 12), (2, 208, 2), (3, 193, 2)):
    d = heads * 64
    qkv = (torch.randn(V * tokens, 3 * d, device="cuda") * 1.5).bfloat16()
    out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(V, heads, tokens, device="cuda")
    gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
    dout = torch.randn(V * tokens, d, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    gu.ok(lib.ttl_op_attention_bwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(dout), gu.ptr(lse), gu.ptr(dqkv), V, tokens, heads, 0.125, gu.stream()))
    torch.cuda.synchronize()
    print("attention fwd+bwd", V, tokens, heads, float(out.float().abs().mean()), float(dqkv.float().abs().mean()))
PY
timeout 900 compute-sanitizer --tool racecheck python /tmp/attn_small.py > $O/racecheck_attention.log 2>&1; tail -5 $O/racecheck_attention.log
timeout 600 compute-sanitizer --tool memcheck python /tmp/attn_small.py > $O/memcheck_attention.log 2>&1; tail -3 $O/memcheck_attention.log
