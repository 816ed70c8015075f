#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for dbg in 5 0; do for u in 0 1 3 4 5; do
CHZ_RING_DBG=$dbg CHZ_RING_UNPACK=$u python tools/exp/bench_paths.py 1024,1,16,16,11 >> $O/r2e_unpack.jsonl 2>>$O/r2e_err.txt
done; done
for dbg in 6 4 8 12 14; do
CHZ_RING_DBG=$dbg CHZ_RING_UNPACK=3 python tools/exp/bench_paths.py 1024,1,16,16,11 >> $O/r2e_unpack.jsonl 2>>$O/r2e_err.txt
done
cat $O/r2e_unpack.jsonl | cut -c1-30,80-260
