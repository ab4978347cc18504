#!/usr/bin/env python3
"""Key per-kernel metrics from an ncu report: python tools/ncu_keys.py X.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
K = [("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"),
     ("warps/SM", "sm__warps_active.avg.per_cycle_active"), ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
     ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
     ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
     ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_active"), ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"), ("winst", "smsp__inst_executed.sum"),
     ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
     ("long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
     ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
     ("math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
     ("mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
     ("lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
     ("short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
     ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")]
for row in r[2:]:
    name = row[h.index("Kernel Name")].split("(")[0]
    vals = []
    for lab, k in K:
        v = row[h.index(k)] if k in h else "-"
        try:
            v = f"{float(v):.4g}"
        except ValueError:
            pass
        vals.append(f"{lab}={v}")
    print(name, " ".join(vals))
