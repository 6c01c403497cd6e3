#!/usr/bin/env python
"""Print the headline numbers and the per-kernel table of bench.py JSON lines (one per log file)."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            print(line[:300].rstrip())
            continue
        d = json.loads(line)
        r = d.get("roofline", {})
        print(f"{f}: {d['value']:.1f} {d['unit']}  {d['ms_per_step']:.3f} ms/step  cycle frac {d.get('cycle_roofline_frac', 0):.3f}  "
              f"e2e {d['e2e']['value']:.1f}  launches {d.get('gpu_launches')}  top {r.get('kernel')} frac {r.get('frac')}")
        for k in d.get("kernels", []):
            print(f"    {k['name']:28s} {k['ms_per_cycle']:.4f} ms  x{k['launches_per_cycle']:.0f}  {k.get('algorithmic_gbs') or 0:.0f} GB/s")
