#!/bin/bash
# round 2, GPU call Q: ncu full captures of the role-split ring kernel (critical and 2x oversampled)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2q_*
export STEPS=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chan_ring_ws -s 2 -c 1 -o $O/r2q_ws_os1 python tools/exp/bench_paths.py 1024,1,16,16,0 > $O/r2q_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chan_ring_ws -s 2 -c 1 -o $O/r2q_ws_os2 python tools/exp/bench_paths.py 1024,2,16,16,0,560000000 > $O/r2q_ncu2.log 2>&1
ls -la $O/r2q_*; tail -n 3 $O/r2q_ncu2.log
