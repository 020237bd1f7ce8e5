"""hottest SASS instructions (by warp-stall samples) of one kernel in an .ncu-rep captured with --import-source on
usage: python tools/ncusrc.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines())]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si, src = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
data = []
for r in rows[hi + 1:]:
    if len(r) <= si or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append((int(r[si] or 0), r[src].strip(), len(data)))
tot = sum(d[0] for d in data) or 1
print(f"{rows[0][1][:100] if rows[0] else ''}: {len(data)} instructions, {tot} samples")
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0] * 100 / tot:5.1f}%  #{d[2]:4d}  {d[1][:120]}")
