# 257-token attention forward with P kept in TMEM (attention_fwd_px_kernel, default) against attention_fwd_pp_kernel<true> (TTL_ATTN=pp):
# kernel tests, ViT-L/14 parity tests, the kernel alone, the ViT-L/14 step (alternating runs on one box).
export PYTHONPATH=.
O=gpurun_out/${SLOT:-s63}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py tests/test_gpu_e2e.py -m gpu -x -q -k "attention or l14 or config4" 2>&1 | tail -5
for m in pp px; do
  TTL_ATTN=$m ATTN_BENCH_FWD=64,576 ATTN_BENCH_TOKENS=257 ATTN_BENCH_HEADS=16 timeout 120 python tools/attn_bench.py 2>&1 | grep "attention fwd" | sed "s/^/$m /"
done
FL="--arch ViT-L/14 --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-torch-baseline --no-live-traffic"
for rep in 1 2; do
  for m in pp px; do
    TTL_ATTN=$m timeout 400 python bench.py $FL > $O/vitl14_${m}_$rep.json 2>>$O/err.log
  done
done
for f in $O/vitl14_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],2), round(d['windows']['median'],2), d['clocks']['sm_mhz'], d['gpu_launches'])"; done
tail -3 $O/err.log
