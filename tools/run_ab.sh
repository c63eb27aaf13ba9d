export PYTHONPATH=.
mkdir -p gpurun_out/s50
TTL_FUSE_LN=1 timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5
FL="--steps 60 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-torch-baseline --no-live-traffic"
for rep in 1 2; do
  for f in 0 1; do
    TTL_FUSE_LN=$f timeout 300 python bench.py $FL > gpurun_out/s50/fuse${f}_$rep.json 2>>gpurun_out/s50/err.log
  done
done
for f in gpurun_out/s50/*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['windows']['median'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"; done
tail -3 gpurun_out/s50/err.log
