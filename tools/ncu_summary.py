"""Summarise an `ncu --set full` report as JSON: one record per captured launch with the counters DESIGN.md quotes.
    python tools/ncu_summary.py gpurun_out/r02_prof_ipa.ncu-rep "<source note>" > profiles/r02_ncu_full_summary.json"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = {"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], "launches": []}
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
    for k in WANT:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i] + " " + units[i]).strip()
    out["launches"].append(d)
print(json.dumps(out, indent=1))
