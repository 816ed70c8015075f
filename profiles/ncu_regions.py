#!/usr/bin/env python
"""Split a kernel's stall samples and executed instructions into the code regions between barriers:
usage: python profiles/ncu_regions.py report.ncu-rep [kernel_index]  -- one line per region (delimited by
BAR.SYNC / SYNCS...TRYWAIT), with samples, instruction counts and the dominant stall reasons."""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for row in src:
    if row and row[0] == "Kernel Name":
        cur = []; blocks.append(cur)
    elif cur is not None:
        cur.append(row)
blk = blocks[kidx]
h = blk[0]; jx = {k: i for i, k in enumerate(h)}
data = [x for x in blk[1:] if len(x) == len(h)]
stall_cols = [k for k in h if k.startswith("stall_")]
regions = []
cur = {"start": 0, "n": 0, "samples": 0, "exec": 0, "ops": Counter(), "st": Counter(), "first": ""}
tot = sum(int(x[jx["# Samples"]]) for x in data)
for i, x in enumerate(data):
    s = x[jx["Source"]].strip()
    op = [t for t in s.split() if not t.startswith("@")]
    op = op[0].split(".")[0] if op else "?"
    cur["n"] += 1
    cur["samples"] += int(x[jx["# Samples"]])
    e = int(x[jx["Instructions Executed"]])
    cur["exec"] += e
    cur["ops"][op] += e
    for c in stall_cols:
        try: cur["st"][c[6:]] += int(x[jx[c]])
        except ValueError: pass
    if s.startswith("BAR.SYNC") or "TRYWAIT" in s or s.startswith("EXIT"):
        cur["end"] = i; cur["delim"] = s[:40]
        regions.append(cur)
        cur = {"start": i + 1, "n": 0, "samples": 0, "exec": 0, "ops": Counter(), "st": Counter()}
if cur["n"]:
    cur["end"] = len(data); cur["delim"] = "end"; regions.append(cur)
print(f"total samples {tot}")
for r in regions:
    if r["samples"] < tot * 0.002: continue
    ops = " ".join(f"{k}:{v // 1000}k" for k, v in r["ops"].most_common(7))
    st = " ".join(f"{k}:{100 * v // max(1, r['samples'])}%" for k, v in r["st"].most_common(5))
    print(f"[{r['start']:5d}-{r['end']:5d}] {100 * r['samples'] / tot:5.1f}% samples  exec {r['exec'] / 1e6:7.1f}M  | {ops} | {st} | ends: {r['delim']}")
