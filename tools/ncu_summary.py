"""Summarise an .ncu-rep (raw + source pages) into a short text report: key metrics, stall reasons, instruction mix and
the hottest SASS lines.  usage: ncu_summary.py <report.ncu-rep> <env_steps_in_launch>"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, env_steps = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
pat = (r"Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__warps_active.avg.pct|"
       r"launch__registers_per_thread$|launch__occupancy_limit_(registers|shared_mem)|smsp__issue_active.avg.pct|"
       r"launch__shared_mem_per_block_dynamic|launch__grid_size|launch__block_size|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|"
       r"smsp__average_warps_issue_stalled_.*_per_issue_active|smsp__inst_executed.sum$|sm__pipe_fma_cycles_active.avg.pct|"
       r"sm__inst_executed_pipe_(fma|alu|lsu|xu).*sum$|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct")
print(f"report: {rep}")
inst = None
for i, k in enumerate(hdr):
    if re.search(pat, k):
        v = vals[i]
        if "stalled" in k:
            try:
                if float(v) < 0.05:
                    continue
            except ValueError:
                pass
        print(f"  {k:88s} {v:>18s} {units[i]}")
        if k == "smsp__inst_executed.sum":
            inst = float(v)
warp_steps = env_steps / 32
if inst:
    print(f"  => warp instructions per warp-env-step: {inst / warp_steps:.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
data = rows[2:]
isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops = collections.Counter()
tots = sum(int(r[ismp]) for r in data)
for r in data:
    parts = r[isrc].split()
    if not parts:
        continue
    op = (parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]
    ops[op] += int(r[iex]) / warp_steps
print("  instruction mix per env-step:", ", ".join(f"{k} {v:.1f}" for k, v in ops.most_common(16)))
print("  hottest SASS lines (stall samples):")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:14]):
    r = data[i]
    print(f"    #{i:5d} exec/step {int(r[iex]) / warp_steps:5.2f} samples {int(r[ismp]) / tots * 100:4.1f}%  {r[isrc].strip()[:70]}")
