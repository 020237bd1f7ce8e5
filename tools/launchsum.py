"""per-kernel mean device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list"""
import csv, sys
from collections import defaultdict
for p in sys.argv[1:]:
    rows = [r for r in csv.reader(open(p)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print(p)
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {sum(v) / tot * 100:5.1f}%  n={len(v):3d}  mean {sum(v) / len(v) / 1e6:9.4f} ms  {k[:110]}")
