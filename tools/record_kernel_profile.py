#!/usr/bin/env python
"""Record ncu-derived constants of the production walker kernel in profiles/walker_inst_per_step.json, keyed by the
hash of the kernel sources that were profiled (bench.kernel_source_hash), so that bench.py only uses them for the
code it is timing.

    python tools/record_kernel_profile.py KEY profiles/<summary>.txt WALKERS [KEY SUMMARY WALKERS ...]

KEY = N50M50g1.0 etc.; SUMMARY = output of profiles/summarize_ncu.py for one launch of WALKERS walkers.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

path = os.path.join(ROOT, "profiles", "walker_inst_per_step.json")
sha = kernel_source_hash()
d = {"source_sha256": sha, "entries": {}}
if os.path.exists(path):
    old = json.load(open(path))
    if old.get("source_sha256") == sha:
        d = old
args = sys.argv[1:]
for key, summary, walkers in zip(args[0::3], args[1::3], args[2::3]):
    txt = open(os.path.join(ROOT, summary)).read()
    inst = float(re.search(r"warp-instructions per 32 units: ([0-9.]+)", txt).group(1))
    rd = re.search(r"dram__bytes_read.sum\s+([0-9.]+) (\w+)", txt)
    wr = re.search(r"dram__bytes_write.sum\s+([0-9.]+) (\w+)", txt)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = float(rd.group(1)) * scale[rd.group(2)] + float(wr.group(1)) * scale[wr.group(2)]
    issue = float(re.search(r"smsp__issue_active.avg.pct_of_peak_sustained_active\s+([0-9.]+)", txt).group(1))
    shp = re.search(r"l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed\s+([0-9.]+)", txt)
    d["entries"][key] = {"inst_per_walker_step": inst, "dram_bytes_per_walker": dram / float(walkers),
                         "ncu_issue_active_pct": issue, "ncu_shared_pipe_pct": float(shp.group(1)) if shp else None,
                         "profile": summary}
d["_note"] = ("written by tools/record_kernel_profile.py from `ncu --set full` captures; bench.py uses an entry only when "
              "source_sha256 equals the hash of the kernel sources it is timing (bench.kernel_source_hash)")
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d, indent=1))
