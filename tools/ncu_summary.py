"""Summarise an `ncu --set full` capture for profiles/: the raw page as CSV plus a small JSON with the numbers bench.py
and DESIGN.md quote (duration, DRAM bytes, pipe utilisation, issue slots, top stall reasons).
  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r2_fused_c4_ncu --cells 125000 [--kernel-index 0]"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
cells = int(sys.argv[sys.argv.index("--cells") + 1]) if "--cells" in sys.argv else None
ki = int(sys.argv[sys.argv.index("--kernel-index") + 1]) if "--kernel-index" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(out + "_raw.csv", "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + ki]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(name, scale_units=True):
    u, v = d[name]
    x = float(v.replace(",", ""))
    if scale_units:
        x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return x


s = {"kernel": d["Kernel Name"][1], "cells": cells, "duration_s": num("gpu__time_duration.sum"),
     "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
     "registers_per_thread": num("launch__registers_per_thread", False),
     "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
     "pipe_tensor_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
     "pipe_xu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", False),
     "pipe_fma_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
     "pipe_alu_pct": num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", False),
     "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active", False),
     "inst_executed": num("smsp__inst_executed.sum", False),
     "stalls_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(v[1]) for k, v in d.items()
                          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}}
s["stalls_per_issue"] = dict(sorted(s["stalls_per_issue"].items(), key=lambda kv: -kv[1])[:8])
json.dump(s, open(out + ".json", "w"), indent=1)
print(json.dumps(s, indent=1))
