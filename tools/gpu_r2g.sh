#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu -k "1024" 2>&1 | tail -2
for dbg in 5 0; do for u in 0 1 3 4; do
CHZ_RING_DBG=$dbg CHZ_RING_UNPACK=$u python tools/exp/bench_paths.py 1024,1,16,16,11 >> $O/r2g.jsonl 2>>$O/r2g_err.txt
done; done
CHZ_RING_UNPACK=1 python tools/exp/bench_paths.py 1024,2,16,16,11 1024,1,12,12,11 >> $O/r2g.jsonl 2>>$O/r2g_err.txt
cat $O/r2g.jsonl | cut -c1-30,80-260
