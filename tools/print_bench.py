"""One-line summary of bench.py JSON lines: python tools/print_bench.py FILE [FILE ...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path).read().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    e2e = (d.get("e2e") or {}).get("value")
    roof = d.get("roofline") or {}
    print(f"{path}: value {d['value']:.1f} {d['unit']} ({d['config']['workload']}), window median {d['windows']['median']:.1f}, "
          f"e2e {e2e if e2e is None else round(e2e, 1)}, SM {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}, "
          f"GEMM frac {roof.get('frac')}, torch baseline {(d.get('torch_gpu_baseline') or {}).get('value')}")
