#!/usr/bin/env python3
"""Summary of every launch in an .ncu-rep (ncu --set full): the metrics DESIGN.md / bench.py quote, as JSON.
usage: summarize.py <report.ncu-rep> [out.json]        (runs in the build container, no GPU needed)"""
import csv, json, subprocess, sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]
STALL = "smsp__average_warps_issue_stalled_"
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.split("\n")))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    o = {k: (d[k] + (" " + u[k] if u.get(k) else "")) for k in KEYS if k in d}
    st = {k[len(STALL):-len("_per_issue_active.ratio")]: float(d[k]) for k in d if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k]}
    tot = sum(st.values()) or 1.0
    o["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot >= 0.01}
    out.append(o)
txt = json.dumps(out, indent=1)
open(sys.argv[2], "w").write(txt) if len(sys.argv) > 2 else print(txt)
