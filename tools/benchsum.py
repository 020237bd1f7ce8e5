"""print a one-line summary of bench.py JSON lines (files given as arguments)"""
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable", e); continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    print(f"{p}: value {d['value']:.4g} q/s  ms/step {d['ms_per_step']:.3f}  kernel_ms {r.get('kernel_ms', 0):.3f}  e2e {e.get('value') or 0:.4g}  "
          f"roofline {r.get('achieved', 0):.0f} GB/s ({r.get('frac', 0):.2f})  launches {d.get('gpu_launches')}  bad {d.get('bad_status')}  "
          f"cpu {c.get('value')} ({c.get('kind')}, {c.get('cores')} cores)  clocks {d.get('clocks')}")
