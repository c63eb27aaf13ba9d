# Final lines of the round (1 x B200): full GPU test suite, smoke, default bench (what the driver runs) + the other BASELINE configs,
# reference arm, ncu --set full of the first four CTA-pair GEMM launches (layer 0 qkv, out-proj, fc1, fc2 at M = 113 472).
export PYTHONPATH=.
O=gpurun_out/${SLOT:-s56}; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_full.log 2>&1; tail -4 $O/pytest_full.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > $O/bench_tpt.json 2> $O/bench.err
python bench.py --head deyo --no-cpu-baseline --no-torch-baseline --no-live-traffic > $O/bench_deyo.json 2>> $O/bench.err
python bench.py --tta-steps 4 --no-cpu-baseline --no-torch-baseline --no-live-traffic > $O/bench_tta4.json 2>> $O/bench.err
python bench.py --classes 200 --no-cpu-baseline --no-torch-baseline --no-live-traffic > $O/bench_c200.json 2>> $O/bench.err
python bench.py --arch ViT-L/14 --steps 40 --warmup 3 --no-cpu-baseline --no-torch-baseline --no-live-traffic > $O/bench_vitl14.json 2>> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2>> $O/bench.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d.get("e2e") or {}
r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value %.1f" % d["value"], "e2e %.1f" % (e.get("value") or 0), "windows", (d.get("windows") or {}).get("median"),
      "roofline", r.get("frac"), "clocks", d.get("clocks"), "torch", (d.get("torch_gpu_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
done
tail -3 $O/bench.err
NOB="--no-e2e --no-roofline --no-cpu-baseline --no-torch-baseline --no-live-traffic"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2_kernel -c 4 -f -o $O/gemm2_full \
  python bench.py --steps 1 --warmup 2 --preheat-s 0 --profile-region $NOB > $O/ncu_gemm2.log 2>&1
ncu -i $O/gemm2_full.ncu-rep --page raw --csv > $O/gemm2_full_raw.csv 2>/dev/null
python tools/ncu_trim.py $O/gemm2_full_raw.csv $O/gemm2_full_trim.csv "ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -c 4, bench.py --steps 1 --warmup 2 --profile-region (9 concurrent samples, M = 113472 rows), final round-2 tree" | head -12
rm -f $O/gemm2_full.ncu-rep
ls -la $O
