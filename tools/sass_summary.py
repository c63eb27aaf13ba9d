#!/usr/bin/env python
"""profiles/*_sass_summary.txt: per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses
(UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit, HMMA = legacy
mma.sync), from `cuobjdump -sass` of the built library.  Usage: python tools/sass_summary.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ttl-test-time-low-rank-adaptation_b200", "ttl_b200", "libttl_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "HMMA", "MUFU.EX2", "LDSM", "SYNCS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::|ttl::|<unnamed>::", "", cur)
            cur = re.sub(r"\(.*$", "", cur)
            order.append(cur)
            continue
        if cur is None:
            continue
        for k in KEYS:
            if re.search(r"\b" + re.escape(k) + r"\b", line) or (k in line and k.endswith("EX2")):
                counts[cur][k] += 1
    lines = [f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)",
             "# " + " ".join(f"{k:>8s}" for k in KEYS) + "  kernel"]
    for fn in order:
        c = counts[fn]
        lines.append("  " + " ".join(f"{c.get(k, 0):8d}" for k in KEYS) + "  " + fn[:110])
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
