#!/usr/bin/env python
"""Turns the ncu artefacts gpurun brought back (gpurun_out/) into the committed summaries:
  python profiles/summarize.py r01      ->  profiles/r01_launches.md, profiles/r01_sweep_metrics.md
Reads gpurun_out/<tag>_launches.csv (ncu --metrics gpu__time_duration.sum --csv) and
gpurun_out/<tag>_sweep.ncu-rep (ncu --set full)."""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
src = os.path.join(ROOT, "gpurun_out")

# ---- launch list ---------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(src, f"{tag}_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(list)
for r in rows[1:]:
    try:
        agg[re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:90]].append(float(r[ix["Metric Value"]]))
    except (ValueError, IndexError):
        pass
unit = rows[1][ix["Metric Unit"]]
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: per-kernel device time (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
    f.write("Command: `python bench.py --steps 1 --warmup 1 --sweeps 20 --no-cpu-baseline` (first 400 launches; "
            "cold-cache, serialised: compare shares, not absolutes).\n\n")
    f.write(f"| kernel | launches | mean ({unit}) | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.0f} | {100*sum(v)/tot:.1f}% |\n")

# ---- full metrics of the sweep kernel --------------------------------------------------------------
rep = os.path.join(src, f"{tag}_sweep.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
idx = {h: i for i, h in enumerate(hdr)}
with open(os.path.join(out_dir, f"{tag}_sweep_metrics.md"), "w") as f:
    f.write(f"# {tag}: ncu --set full of the sweep kernel (one capture per distinct kernel)\n\n")
    f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:k_sweep_rows -s 4 -c 4 "
            "python profiles/scripts/prof_c3.py 7 6` on one B200.  ncu flushes caches "
            "between replays, so `dram__bytes_read` is the cold-cache figure (whole state + couplings); in the "
            "benchmark loop the 32 MiB state stays in L2.\n\n")
    seen = set()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in want:
            if w in idx:
                f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
        f.write("\n")
print("wrote", tag)
