#!/usr/bin/env python
"""Per-region dynamic instruction counts from an ncu report's SASS source page.

    python tools/ncu_regions.py report.ncu-rep <units> [start:end:name ...]   (offsets in hex, from the kernel start)

Without regions prints the loop structure (backward branches) and the hottest instructions.
"""
import csv
import subprocess
import sys
from collections import Counter

ALU = {"LOP3", "SHF", "SEL", "ISETP", "IADD3", "VIADD", "FSEL", "PRMT", "VIMNMX", "VIADDMNMX", "LEA", "PLOP3", "MOV",
       "FSETP", "FLO", "BREV", "IABS", "FMNMX", "SGXT", "BMSK", "LOP", "IADD", "I2FP", "F2FP"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "IDP", "HFMA2", "HADD2", "HMUL2"}
FP64 = {"DFMA", "DADD", "DMUL", "DSETP"}
XU = {"POPC", "MUFU", "I2F", "F2I", "F2F", "FRND"}

rep, units = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
ins = []
base = None
for r in rows[2:]:
    a = int(r[ia], 16)
    base = a if base is None else base
    toks = [t for t in r[isrc].split() if not t.startswith("@")]
    ins.append((a - base, toks[0].split(".")[0], int(r[iex]), r[isrc].strip()))
warps = units / 32.0


def pipe(op):
    return "alu" if op in ALU else "fma" if op in FMA else "fp64" if op in FP64 else "xu" if op in XU else "other"


regions = []
for s in sys.argv[3:]:
    a, b, name = s.split(":")
    regions.append((int(a, 16), int(b, 16), name))
if not regions:
    regions = [(0, 1 << 30, "all")]
for a, b, name in regions:
    tot = Counter()
    ops = Counter()
    for off, op, ex, _ in ins:
        if a <= off < b:
            tot[pipe(op)] += ex
            ops[op] += ex
    n = sum(tot.values())
    print(f"{name:24s} {n / warps:8.2f} inst/unit  " + "  ".join(f"{k}={v / warps:.1f}" for k, v in tot.most_common()))
    print("    " + "  ".join(f"{k}:{v / warps:.1f}" for k, v in ops.most_common(12)))
