#!/usr/bin/env python3
"""Text summary of an `ncu --set full` capture for profiles/: per captured kernel the duration, occupancy, pipe
utilisation, DRAM bytes, instruction count and the top stall reasons (warps stalled per issue).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "what was captured" > profiles/x_ncu.txt
"""
import csv
import io
import subprocess
import sys

rep, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg"]
stalls = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
units = dict(zip(hdr, rows[1]))
print("ncu --set full --clock-control none; capture %s (not committed)" % rep)
if note:
    print(note)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print()
    print("Kernel Name".ljust(96), d.get("Kernel Name", "")[:110])
    for w in WANT:
        if d.get(w) not in (None, ""):
            print(w.ljust(96), d[w], units.get(w, ""))
    top = sorted(((float(d[h]), h) for h in stalls if d.get(h) not in (None, "")), reverse=True)[:6]
    for v, h in top:
        print(("stall " + h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")).ljust(96),
              "%.2f warps per issue" % v)
