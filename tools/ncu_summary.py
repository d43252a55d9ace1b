#!/usr/bin/env python
"""Summarise an ncu report (raw page) and a launch list into the text files kept under profiles/.
   python tools/ncu_summary.py <report.ncu-rep> <launches.csv> <out_prefix> "<title>" """
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, launches, out, title = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
pat = re.compile(r'^(gpu__time_duration.sum|sm__inst_executed_pipe_(alu|fma|fmaheavy|fmalite|lsu|cbu|xu|uniform)\.avg\.pct_of_peak_sustained_active|'
                 r'smsp__issue_active.avg.pct_of_peak_sustained_active|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|'
                 r'launch__grid_size|launch__block_size|launch__occupancy_limit_(registers|shared_mem|warps)|launch__shared_mem_per_block_dynamic|'
                 r'dram__bytes_(read|write).sum|smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active.ratio|smsp__inst_executed.sum|'
                 r'sm__cycles_elapsed.avg|smsp__warps_eligible.avg.per_cycle_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|'
                 r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|'
                 r'smsp__thread_inst_executed_per_inst_executed.ratio|sm__sass_thread_inst_executed_op_(ffma|fmul|fadd)_pred_on.sum|'
                 r'smsp__sass_thread_inst_executed_op_(ffma|fmul|fadd)_pred_on.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed)$')
with open(out + "_ncu_summary.txt", "w") as f:
    f.write(f"# {title}\n# source: {rep} ({len(rows) - 2} launch(es) captured; one column per launch)\n")
    ik = hdr.index("Kernel Name")
    f.write("# kernels: " + " | ".join(r[ik][:90] for r in rows[2:]) + "\n")
    for i, (h, u) in enumerate(zip(hdr, units)):
        if pat.search(h):
            f.write(f"{h:95s} " + "  ".join(f"{r[i]:>18s}" for r in rows[2:]) + f"  {u}\n")
rows = list(csv.reader(open(launches)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) > iv:
        n = r[ik].split("(")[0][:80]
        d[n][0] += 1
        d[n][1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in d.values())
with open(out + "_launch_shares.txt", "w") as f:
    f.write(f"# {title}\n# per-kernel device time from the ncu launch list {launches} (cold-cache, serialised: compare shares)\n")
    for n, v in sorted(d.items(), key=lambda x: -x[1][1]):
        f.write(f"{n:82s} launches {v[0]:4d}  time {v[1] / 1e6:10.3f} ms  share {v[1] / tot:.4f}\n")
print(open(out + "_ncu_summary.txt").read())
print(open(out + "_launch_shares.txt").read())
