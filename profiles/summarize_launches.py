"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (times are cold-cache and
serialised: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'total us':>12} {'n':>5} {'share':>6} {'avg us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {v[0]:5d} {100 * v[1] / tot:5.1f}% {v[1] / v[0]:9.1f}  {k[:130]}")
    print(f"total {tot:.1f} us over {len(rows)} launches")


if __name__ == "__main__":
    main(sys.argv[1])
