"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  Usage: python tools/launch_summary.py launches.csv > launches.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    with open(path) as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    kn, mv, mu = ix["Kernel Name"], ix["Metric Value"], ix["Metric Unit"]
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        us = v / 1e3 if r[mu] in ("ns", "nsecond") else (v if r[mu] in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r[kn])[:150]
        t, n = agg.get(name, (0.0, 0))
        agg[name] = (t + us, n + 1)
    total = sum(t for t, _ in agg.values())
    print("    total us     n  share    avg us  kernel")
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t:12.1f} {n:5d} {100 * t / total:5.1f}% {t / n:9.1f}  {name}")
    print(f"total {total:.1f} us over {sum(n for _, n in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
