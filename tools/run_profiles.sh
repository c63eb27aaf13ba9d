# Profiles of the 257-token attention forward with P in TMEM (attention_fwd_px_kernel) on the final tree: ncu --set full, launch list
# of a ViT-L/14 step, sanitizers over the kernel, the ViT-L/14 bench line, the full GPU test suite.
export PYTHONPATH=.
O=gpurun_out/${SLOT:-s64}; mkdir -p $O
NOB="--no-e2e --no-roofline --no-cpu-baseline --no-torch-baseline --no-live-traffic"
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_full.log 2>&1; grep -E "passed|failed" $O/pytest_full.log
ncu --set full --clock-control none --import-source on -k regex:attention_fwd_px -c 2 -f -o $O/attn_px \
  env ATTN_BENCH_FWD=576 ATTN_BENCH_TOKENS=257 ATTN_BENCH_HEADS=16 python tools/attn_bench.py > $O/n1.log 2>&1
ncu -i $O/attn_px.ncu-rep --page raw --csv > $O/attn_px_raw.csv 2>/dev/null
python tools/ncu_trim.py $O/attn_px_raw.csv $O/attn_px_trim.csv "ncu --set full --clock-control none --import-source on -k regex:attention_fwd_px -c 2, tools/attn_bench.py at 576 views x 16 heads x 257 tokens (ViT-L/14), final round-2 tree" | head -6
rm -f $O/attn_px.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_vitl14.csv \
  python bench.py --arch ViT-L/14 --steps 2 --warmup 2 --preheat-s 0 --profile-region $NOB > $O/l1.log 2>&1
python tools/launch_summary.py $O/launches_vitl14.csv > $O/launches_vitl14.txt; head -8 $O/launches_vitl14.txt
cat > /tmp/attn257.py <<'PY'
import os, sys
sys.path[:0] = [".", "ttl-test-time-low-rank-adaptation_b200", "tests"]
import torch, gpu_util as gu
lib = gu.lib()
for (V, tokens, heads) in ((20, 257, 16), (3, 257, 2)):
    d = heads * 64
    qkv = (torch.randn(V * tokens, 3 * d, device="cuda") * 1.5).bfloat16()
    out = torch.empty(V * tokens, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(V, heads, tokens, device="cuda")
    gu.ok(lib.ttl_op_attention_fwd(gu.ptr(qkv), gu.ptr(out), gu.ptr(lse), V, tokens, heads, 0.125, gu.stream()))
    torch.cuda.synchronize()
    print("attention fwd", V, tokens, heads, float(out.float().abs().mean()), float(lse.mean()))
PY
timeout 900 compute-sanitizer --tool memcheck python /tmp/attn257.py > $O/memcheck_attention_px.log 2>&1; tail -3 $O/memcheck_attention_px.log
timeout 900 compute-sanitizer --tool racecheck python /tmp/attn257.py > $O/racecheck_attention_px.log 2>&1; tail -3 $O/racecheck_attention_px.log
timeout 900 compute-sanitizer --tool synccheck python /tmp/attn257.py > $O/synccheck_attention_px.log 2>&1; tail -3 $O/synccheck_attention_px.log
python bench.py --arch ViT-L/14 --steps 40 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-live-traffic > $O/bench_vitl14.json 2> $O/bench.err
python tools/print_bench.py $O/bench_vitl14.json | cut -c1-300
ls -la $O
