#!/usr/bin/env python
"""profiles/gemm2_traffic.json (+ a trimmed CSV) from one `ncu --set full` capture of the first four CTA-pair GEMM launches of a
bench step (qkv, out-proj, fc1, fc2 of encoder layer 0):

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2_kernel -c 4 \
        -o gpurun_out/gemm2_full -f python bench.py --steps 1 --warmup 2 --profile-region --no-e2e --no-roofline --no-cpu-baseline
    ncu -i gpurun_out/gemm2_full.ncu-rep --page raw --csv > /tmp/gemm2_raw.csv
    python tools/make_traffic_json.py /tmp/gemm2_raw.csv --samples 9 --tag s9 --source "gpurun sNN"

bench.py reads the JSON for `roofline.traffic` (DRAM bytes per launch of the dominant kernel)."""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[unit]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--samples", type=int, required=True, help="concurrent samples of the captured bench step")
    ap.add_argument("--views", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=197)
    ap.add_argument("--tag", default="s9")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    with open(a.raw_csv) as f:
        rows = [r for r in csv.reader(l for l in f if not l.startswith("=="))]
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    M, d, F = a.samples * a.views * a.tokens, 768, 3072
    mb = lambda n: n / 1e6
    # the four launches in issue order: qkv, out-proj, fc1, fc2
    shapes = [("qkv", 3 * d, d, "bias->bf16", mb(M * d * 2) + mb(3 * d * d * 2), mb(M * 3 * d * 2)),
              ("out-proj", d, d, "bias+residual->f32", mb(M * d * 2) + mb(M * d * 4) + mb(d * d * 2), mb(M * d * 4)),
              ("fc1", F, d, "bias+QuickGELU->bf16", mb(M * d * 2) + mb(F * d * 2), mb(M * F * 2)),
              ("fc2", d, F, "bias+residual->f32", mb(M * F * 2) + mb(M * d * 4) + mb(d * F * 2), mb(M * d * 4))]
    assert len(data) >= 4, "expected the first four gemm2_kernel launches of the profiled step"
    per, trimmed = [], []
    for (nm, N, K, epi, rd, wr), r in zip(shapes, data[:4]):
        g = lambda k: r[ix[k]]
        u = lambda k: units[ix[k]]
        dr, dw = to_mb(g("dram__bytes_read.sum"), u("dram__bytes_read.sum")), to_mb(g("dram__bytes_write.sum"), u("dram__bytes_write.sum"))
        per.append({"kernel": r[ix["Kernel Name"]][:64], "what": f"{nm}  M={M} N={N} K={K}  {epi}: algorithmic {rd:.1f} MB read, {wr:.1f} MB write",
                    "algorithmic_MB": round(rd + wr, 1), "dram_MB": round(dr + dw, 1), "dram_read_MB": round(dr, 1), "dram_write_MB": round(dw, 1),
                    "us_under_ncu": round(to_us(g("gpu__time_duration.sum"), u("gpu__time_duration.sum")), 2),
                    "tensor_pipe_active_pct": round(float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")), 1),
                    "dram_pct_of_peak": round(float(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")), 1),
                    "l2_hit_pct": round(float(g("lts__t_sector_hit_rate.pct")), 1)})
        trimmed.append([r[ix["Kernel Name"]]] + [g(c) for c in COLS])
    src = (f"ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -c 4, bench.py --steps 1 --warmup 2 --profile-region "
           f"({a.samples} concurrent samples, M = {M} rows), {a.source}; profiles/r1_ncu_full_gemm2_{a.tag}.csv")
    out = {"source": src, "unit": "MB per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
           "bytes_per_launch_avg": int(sum(p["dram_MB"] for p in per) / len(per) * 1e6),
           "note": "one launch of each big shape of an encoder layer (layer 0: qkv, out-proj, fc1, fc2); compare dram_MB with algorithmic_MB: "
                   "traffic above the algorithmic bytes would mean wasted re-reads",
           "per_launch": per}
    with open(os.path.join(ROOT, "profiles", "gemm2_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "profiles", f"r1_ncu_full_gemm2_{a.tag}.csv"), "w") as f:
        f.write(f'"# {src}"\n')
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + COLS)
        w.writerow([""] + [units[ix[c]] for c in COLS])
        w.writerows(trimmed)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
