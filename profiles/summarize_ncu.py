#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box): headline metrics, stall breakdown, dynamic
opcode mix and the hottest SASS lines of the first captured kernel.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [kernel_index] > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, rows = raw[0], raw[1], raw[2:]
r = rows[kidx]
ix = {h: i for i, h in enumerate(hdr)}
print(f"report: {rep}   kernels captured: {len(rows)}   showing #{kidx}")
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
for k in keys:
    if k in ix:
        print(f"  {k:75s} {r[ix[k]]} {units[ix[k]]}")
st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for h, i in ix.items()
      if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i]}
tot = sum(st.values()) or 1
print("\nwarp stall samples (all):")
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {k:24s} {v:9.0f} {100 * v / tot:5.1f}%")

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
blocks, cur = [], None
for row in src:
    if row and row[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(row)
blk = blocks[kidx]
h2 = blk[0]
jx = {h: i for i, h in enumerate(h2)}
data = [x for x in blk[1:] if len(x) == len(h2)]
ti = sum(int(x[jx["Instructions Executed"]]) for x in data)
ts = sum(int(x[jx["# Samples"]]) for x in data)
ops, smp = Counter(), Counter()
for x in data:
    toks = [t for t in x[jx["Source"]].split() if not t.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    ops[op] += int(x[jx["Instructions Executed"]])
    smp[op] += int(x[jx["# Samples"]])
print(f"\ndynamic opcode mix (warp instructions executed: {ti}, stall samples: {ts}):")
for op, n in ops.most_common(18):
    print(f"  {op:8s} {n:11d} {100 * n / ti:5.1f}%   samples {100 * smp[op] / ts:5.1f}%")
print("\nhottest SASS lines by stall samples:")
for x in sorted(data, key=lambda x: -int(x[jx["# Samples"]]))[:25]:
    print(f"  {x[jx['# Samples']]:>6} exec={x[jx['Instructions Executed']]:>9} long_sb={x[jx['stall_long_sb']]:>5} "
          f"wait={x[jx['stall_wait']]:>5} short_sb={x[jx['stall_short_sb']]:>5} math={x[jx['stall_math']]:>4} "
          f"bar={x[jx['stall_barrier']]:>4} smem_excess={x[jx['L1 Wavefronts Shared Excessive']]:>8}  {x[jx['Source']].strip()[:80]}")
