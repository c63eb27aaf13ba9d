export PYTHONPATH=.
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "attention" 2>&1 | tail -2
for var in 0 1 2; do echo "variant $var"; TTL_PT_VARIANT=$var ATTN_BENCH_FWD=192,576 timeout 120 python tools/attn_bench.py; done
echo pp; TTL_ATTN=pp ATTN_BENCH_FWD=192,576 timeout 120 python tools/attn_bench.py
TTL_ATTN_DBG=1 ATTN_BENCH_FWD=192 timeout 120 python tools/attn_bench.py 2>&1 | grep "unit [345]"
