"""Compare two `ncu --metrics gpu__time_duration.sum,dram__bytes_*,lts__t_sector_hit_rate.pct --csv` captures of the same step, one
with TTL_FUSE_LN=0 and one with TTL_FUSE_LN=1 (tools/run_fuse_ln.sh): per kernel and grid shape, launches, average duration, DRAM
bytes per launch and L2 hit rate.  Usage: python tools/fuse_ln_summary.py fuse0.csv fuse1.csv"""
import csv
import re
import sys
from collections import OrderedDict

UNIT = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}


def load(path):
    with open(path) as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    ix = {h: i for i, h in enumerate(rows[0])}
    per_launch = OrderedDict()
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        key = r[ix["ID"]]
        d = per_launch.setdefault(key, {"name": re.sub(r"\(.*", "", r[ix["Kernel Name"]]).split("::")[-1]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
    return list(per_launch.values())


def classify(d):
    """The residual-epilogue pair kernel serves out-proj (K = 768) and fc2 (K = 3072): tell them apart by duration."""
    n = d["name"]
    if n.startswith("gemm2_kernel<256, 2, 2>"):
        return n + (" [fc2]" if d["gpu__time_duration.sum"] > 170.0 else " [out-proj]")
    return n


def table(launches):
    agg = OrderedDict()
    for d in launches:
        k = classify(d)
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["gpu__time_duration.sum"]
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
        a[4] += d.get("lts__t_sector_hit_rate.pct", 0.0)
    return agg


def main(p0, p1):
    for label, p in (("TTL_FUSE_LN=0", p0), ("TTL_FUSE_LN=1", p1)):
        agg = table(load(p))
        print(label)
        print("      n   avg us  total us  DRAM rd MB  DRAM wr MB  L2 hit %  kernel")
        tot = 0.0
        for k, (n, t, rd, wr, hit) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            tot += t
            print(f"  {n:5d} {t / n:8.1f} {t:9.1f} {rd / n / 1e6:11.1f} {wr / n / 1e6:11.1f} {hit / n:9.1f}  {k[:90]}")
        print(f"  total {tot:.1f} us (cold-cache, serialised launches: shares, not absolutes)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
