#!/usr/bin/env python3
"""Executed warp instructions of one kernel of an ncu report, binned by SASS index with the opcode mix per bin.
usage: ncu_bins.py report.ncu-rep launch_index [bin]"""
import csv, io, subprocess, sys, collections
rep, k = sys.argv[1], int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 100
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = txt.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[1:]:
    if len(r) != len(hdr) or r[0] == hdr[0]:
        if data: break
        continue
    data.append(r)
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("instructions", len(data), "executed", tot)
for b0 in range(0, len(data), B):
    chunk = data[b0:b0 + B]
    ex = sum(int(r[ix["Instructions Executed"]]) for r in chunk)
    smp = sum(int(r[ix["# Samples"]]) for r in chunk)
    ops = collections.Counter()
    for r in chunk:
        op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
        ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]])
    top = ", ".join(f"{o}:{100*c/max(ex,1):.0f}%" for o, c in ops.most_common(6))
    print(f"[{b0:5d}..{b0+len(chunk)-1:5d}] exec {100*ex/tot:5.1f}%  samples {smp:6d}  {top}")
