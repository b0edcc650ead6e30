#!/usr/bin/env python3
"""Hot spots of one kernel from an ncu report (source page, SASS): instructions with the most stall samples, grouped into regions
between barriers.  usage: ncu_hot.py report.ncu-rep launch_index [top]"""
import csv
import io
import subprocess
import sys

rep, k = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = txt.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[1:]:
    if len(r) != len(hdr) or r[0] == hdr[0]:
        if data:
            break          # a second table (another view of the same kernel) follows
        continue
    data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data)
texec = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("total samples", tot, "warp instructions executed", texec)
# regions between BAR instructions
reg, cur, start = [], 0, 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
acc = {h: 0 for h in stall_cols}
nexec = 0
for i, r in enumerate(data):
    cur += int(r[ix["# Samples"]])
    nexec += int(r[ix["Instructions Executed"]])
    for h in stall_cols:
        acc[h] += int(r[ix[h]])
    if "BAR.SYNC" in r[ix["Source"]] or i == len(data) - 1:
        topst = sorted(acc.items(), key=lambda kv: -kv[1])[:4]
        reg.append((start, i, cur, nexec, topst))
        cur, start, nexec = 0, i + 1, 0
        acc = {h: 0 for h in stall_cols}
print("regions between barriers: [first..last instr] samples (share) executed  top stalls")
for a, b, c, n, st in reg:
    if c > 0.005 * tot:
        print(f"  [{a:5d}..{b:5d}] {c:7d} ({100.0 * c / tot:5.1f}%) exec {n:10d}  " + ", ".join(f"{h[6:]}:{v}" for h, v in st))
print("hottest instructions:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    i = data.index(r)
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"  #{i:5d} {int(r[ix['# Samples']]):6d} ({100.0 * int(r[ix['# Samples']]) / tot:4.1f}%)  {r[ix['Source']].strip()[:70]:70s} {st}")
