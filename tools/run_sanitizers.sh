# compute-sanitizer on the final round-2 tree: memcheck / synccheck over the end-to-end driver, memcheck over the opt-in fused-LayerNorm
# GEMM path (TTL_FUSE_LN, all three settings through its agreement test).
export PYTHONPATH=.
O=gpurun_out/${SLOT:-s57}; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck python tests/sanitizer_driver_b16.py > $O/memcheck.log 2>&1; tail -3 $O/memcheck.log
timeout 900 compute-sanitizer --tool synccheck python tests/sanitizer_driver_b16.py > $O/synccheck.log 2>&1; tail -3 $O/synccheck.log
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_e2e.py -m gpu -x -q -k fused_layernorm > $O/memcheck_fused_ln.log 2>&1; tail -4 $O/memcheck_fused_ln.log
