# Final bench lines of the round (1 x B200): default flags (what the driver runs) + the other BASELINE configs.
export PYTHONPATH=.
O=gpurun_out/s44; mkdir -p $O
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
