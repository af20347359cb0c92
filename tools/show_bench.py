#!/usr/bin/env python
"""Print the headline numbers and the per-kernel-class breakdown of a bench.py JSON line."""
import json
import sys

for path in sys.argv[1:]:
    try:
        b = json.load(open(path))
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    print(f"{path}: value {b['value']:.0f} {b['unit']}  ms/step {b['ms_per_step']:.1f}  e2e {b['e2e']['value']:.0f}  "
          f"launches {b.get('gpu_launches')}  clocks {b.get('clocks')}")
    r = b.get("roofline") or {}
    print(f"  roofline: {r.get('kernel')}: {r.get('achieved', 0):.1f} {r.get('unit')} = {r.get('frac', 0):.3f} of {r.get('peak')}")
    for k, v in (b.get("kernels") or {}).items():
        extra = f"  {v['achieved']:.1f} {v['unit']} ({v['frac']:.3f})" if "achieved" in v else ""
        print(f"  {k:12s} launches {v['launches']:6d}  total {v['ms_total']:8.2f} ms  avg {1e3 * v['ms_total'] / max(v['launches'], 1):8.1f} us{extra}")
    if b.get("cpu_baseline"):
        print("  cpu_baseline:", b["cpu_baseline"]["value"], b["cpu_baseline"]["unit"], "cores", b["cpu_baseline"]["cores"])
