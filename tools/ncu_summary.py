#!/usr/bin/env python3
"""`ncu --set full` report -> one CSV row per captured launch with the metrics DESIGN.md / profiles/README.md quote.
Usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, io, subprocess, sys

rep, dst = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
want = [
    ("time", "gpu__time_duration.sum"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"), ("l1_hit", "l1tex__t_sector_hit_rate.pct"), ("l2_hit", "lts__t_sector_hit_rate.pct"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("lanes/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("branch_uniform_pct", "smsp__sass_average_branch_targets_threads_uniform.pct"), ("occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("warp_inst", "smsp__inst_executed.sum"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_no_inst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
]
have = [(a, m) for a, m in want if m in ix]
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel"] + [a for a, _ in have] + ["units: " + "; ".join(f"{a}={units[ix[m]]}" for a, m in have)])
    for i, r in enumerate(data):
        name = r[ix["Kernel Name"]]
        w.writerow([i, name] + [r[ix[m]] for _, m in have])
print(open(dst).read())
