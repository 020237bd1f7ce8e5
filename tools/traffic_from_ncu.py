"""record `dram__bytes_read.sum + dram__bytes_write.sum` per query of one kernel launch in profiles/traffic.json, keyed by the
source digest of the library the capture was taken from (bench.py refuses an entry whose digest differs from the built library)

usage: python tools/traffic_from_ncu.py report.ncu-rep kernel-regex workload kernel-name queries-per-launch [note]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locointerpolate_b200 import build as product_build  # noqa: E402

rep, pat, workload, kernel, qpl = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
note = sys.argv[6] if len(sys.argv) > 6 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
r = rows[2]
tot = 0.0
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    i = hdr.index(k)
    tot += float(r[i].replace(",", "")) * scale[units[i]]
path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(path)) if os.path.exists(path) else {}
tj[workload] = {"kernel": kernel, "dram_bytes_per_query": tot / qpl, "dram_bytes_per_launch": tot, "queries_per_launch": qpl,
                "source_digest": product_build.source_digest()[:16], "capture": os.path.basename(rep) + (" " + note if note else "")}
json.dump(tj, open(path, "w"), indent=1)
print(workload, kernel, tot / qpl, "B/query", tj[workload]["source_digest"])
