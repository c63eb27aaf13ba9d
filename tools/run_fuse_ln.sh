# Strip-form fc2 -> LayerNorm1 fusion (TTL_FUSE_LN=1): parity, then ncu of the residual-epilogue GEMM launches with and without it
# (duration, DRAM bytes, L2 hit rate), then the launch lists.  Negative result, DESIGN 4.4.
export PYTHONPATH=.
O=gpurun_out/s51; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k "fused_layernorm or zigzag" 2>&1 | tail -5
NOB="--no-e2e --no-roofline --no-cpu-baseline --no-torch-baseline --no-live-traffic"
for f in 0 1; do
  TTL_FUSE_LN=$f timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none --profile-from-start off -k regex:"gemm2_kernel|layernorm_kernel" -c 160 --csv --log-file $O/fuse${f}_gemm2.csv \
    python bench.py --steps 1 --warmup 2 --preheat-s 0 --profile-region $NOB > $O/n$f.log 2>&1
done
python tools/fuse_ln_summary.py $O/fuse0_gemm2.csv $O/fuse1_gemm2.csv | tee $O/summary.txt
