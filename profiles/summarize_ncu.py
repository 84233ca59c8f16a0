#!/usr/bin/env python
"""Summarise an ncu report (one kernel) into the text files kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/walker_r01.ncu-rep <walker-steps in the launch> > profiles/<name>.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None  # walker-steps (or states) processed by the launch

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
val = dict(zip(h, v))
unit = dict(zip(h, u))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg",
    # shared-memory data pipe: what bounds a kernel whose working set lives in shared memory
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
print(f"# ncu summary of {rep}")
print(f"kernel: {val.get('Kernel Name', '?')}")
for k in KEYS:
    if k in val:
        print(f"{k:75s} {val[k]:>18s} {unit[k]}")
print("\n# warp stall reasons (warps per issue-active cycle)")
for k in sorted(h):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        print(f"{k[34:-23]:30s} {float(val[k]):8.3f}")
inst = float(val["smsp__inst_executed.sum"].replace(",", ""))
if units:
    print(f"\nunits in this launch: {units:.0f}; warp-instructions per 32 units: {inst / (units / 32):.1f} "
          f"(= SASS instructions per unit per thread)")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh = rows[1]
ia, ie = hh.index("Source"), hh.index("Instructions Executed")
ops = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ie:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
    op = (m.group(2) if m else r[ia]).split(".")[0]
    n = int(r[ie] or 0)
    ops[op] += n
    tot += n
print("\n# SASS opcode mix (executed warp-instructions)")
for k, n in ops.most_common(25):
    per = f"{n / (units / 32):8.2f}/unit" if units else ""
    print(f"{k:12s} {100 * n / tot:5.1f}%  {per}")
