#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
tools/ubench/pipes > $O/r2d_pipes.txt 2>&1
for dbg in 0 1 2; do for u in 0 1; do
CHZ_RING_DBG=$dbg CHZ_RING_UNPACK=$u python tools/exp/bench_paths.py 1024,1,16,16,11 1024,2,16,16,11 >> $O/r2d_phases.jsonl 2>>$O/r2d_err.txt
done; done
cat $O/r2d_pipes.txt; cat $O/r2d_phases.jsonl
