#!/usr/bin/env python
"""Hottest CUDA source lines of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py report.ncu-rep <units> [top]
"""
import csv
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
fname = None
agg = {}
hdr = None
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iex = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= iex or r[2] != "-":  # keep the per-line aggregate rows only
        continue
    try:
        ex = int(r[iex])
    except ValueError:
        continue
    key = (fname, int(r[0]), r[1].strip()[:100])
    agg[key] = agg.get(key, 0) + ex
tot = sum(agg.values())
print(f"total {tot / units:.1f} inst/unit")
for (f, ln, src), ex in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{ex / units:8.1f}  {f}:{ln:<4d} {src}")
