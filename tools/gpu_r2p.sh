#!/bin/bash
# round 2, GPU call P: role times of the warp-specialised ring kernel (dbg 1 = no FFT, 2 = no FIR, 4 = no loads)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2p_*
for d in 0 1 2 3 4 5 6 7; do
  CHZ_RING_VARIANT=2 CHZ_RING_DBG=$d timeout 300 python tools/exp/bench_paths.py 1024,1,16,16,0 1024,2,16,16,0 >> $O/r2p_bench.jsonl 2>> $O/r2p_bench.err
done
cat $O/r2p_bench.jsonl; tail -n 5 $O/r2p_bench.err
