# Same-box A/B of the fused LayerNorm variants (TTL_FUSE_LN: 0 = off, 1 = fc2 -> LN1, 2 = out-proj -> LN2, 3 = both), alternating runs,
# then ncu (duration, DRAM bytes, L2 hit rate) of the residual-epilogue GEMM launches with the fusion on.
export PYTHONPATH=.
O=gpurun_out/${SLOT:-s55}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "fused_layernorm or zigzag" 2>&1 | tail -15
FL="--steps 60 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline --no-torch-baseline --no-live-traffic"
for rep in 1 2; do
  for f in ${VARIANTS:-0 1 3 2}; do
    TTL_FUSE_LN=$f timeout 300 python bench.py $FL > $O/fuse${f}_$rep.json 2>>$O/err.log
  done
done
for f in $O/fuse*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['windows']['median'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"; done
tail -3 $O/err.log
NOB="--no-e2e --no-roofline --no-cpu-baseline --no-torch-baseline --no-live-traffic"
for f in ${NCU_VARIANTS:-0 1}; do
  TTL_FUSE_LN=$f timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none --profile-from-start off -k regex:"gemm2_kernel|layernorm_kernel" -c 160 --csv --log-file $O/fuse${f}_gemm2.csv \
    python bench.py --steps 1 --warmup 2 --preheat-s 0 --profile-region $NOB > $O/n$f.log 2>&1
done
set -- ${NCU_VARIANTS:-0 1}
python tools/fuse_ln_summary.py $O/fuse$1_gemm2.csv $O/fuse$2_gemm2.csv | tee $O/summary.txt
