export PYTHONPATH=.
mkdir -p gpurun_out/s40
FL="--steps 60 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-torch-baseline --no-live-traffic"
for rep in 1 2; do
  python bench.py $FL > gpurun_out/s40/pt_$rep.json 2>gpurun_out/s40/err.log
  TTL_ATTN=pp python bench.py $FL > gpurun_out/s40/pp_$rep.json 2>>gpurun_out/s40/err.log
done
python bench.py $FL --head deyo > gpurun_out/s40/pt_deyo.json 2>>gpurun_out/s40/err.log
TTL_ATTN=pp python bench.py $FL --head deyo > gpurun_out/s40/pp_deyo.json 2>>gpurun_out/s40/err.log
for f in gpurun_out/s40/*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d.get('windows'), d['clocks'])"; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
