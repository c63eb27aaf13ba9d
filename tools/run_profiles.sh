# Round-2 final profiles: launch lists of the timed region (2 steps of 9 samples) + ncu --set full of the attention forward.
export PYTHONPATH=.
O=gpurun_out/s42; mkdir -p $O
NOB="--no-e2e --no-roofline --no-cpu-baseline --no-torch-baseline --no-live-traffic"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_tpt.csv \
  python bench.py --steps 2 --warmup 2 --preheat-s 0 --profile-region $NOB > $O/l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_deyo.csv \
  python bench.py --steps 2 --warmup 2 --preheat-s 0 --profile-region --head deyo $NOB > $O/l2.log 2>&1
python tools/launch_summary.py $O/launches_tpt.csv > $O/launches_tpt.txt; head -12 $O/launches_tpt.txt
python tools/launch_summary.py $O/launches_deyo.csv > $O/launches_deyo.txt; head -8 $O/launches_deyo.txt
ncu --set full --clock-control none --import-source on -k regex:attention_fwd_pt -c 2 -f -o $O/attn_pt \
  env ATTN_BENCH_FWD=576 python tools/attn_bench.py > $O/n1.log 2>&1
ncu -i $O/attn_pt.ncu-rep --page raw --csv > $O/attn_pt_raw.csv 2>/dev/null
ls -la $O
