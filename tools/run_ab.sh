# Same-box A/B of the fused LayerNorm variants (TTL_FUSE_LN: 0 = off, 1 = fc2 -> LN1, 2 = out-proj -> LN2, 3 = both), alternating runs.
export PYTHONPATH=.
O=gpurun_out/s52; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "fused_layernorm or zigzag" 2>&1 | tail -15
FL="--steps 60 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-torch-baseline --no-live-traffic"
for rep in 1 2; do
  for f in 0 1 3 2; do
    TTL_FUSE_LN=$f timeout 300 python bench.py $FL > $O/fuse${f}_$rep.json 2>>$O/err.log
  done
done
for f in $O/fuse*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['windows']['median'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"; done
tail -3 $O/err.log
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > $O/pytest_full.log 2>&1; tail -30 $O/pytest_full.log
