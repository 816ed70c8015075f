#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu -k "1024" 2>&1 | tail -2
for sc in 0 1; do for u in 0 1; do
CHZ_RING_SCALAR=$sc CHZ_RING_UNPACK=$u python tools/exp/bench_paths.py 1024,1,16,16,11 1024,2,16,16,11 >> $O/r2h.jsonl 2>>$O/r2h_err.txt
done; done
CHZ_RING_UNPACK=1 python tools/exp/bench_paths.py 1024,1,12,12,11 1024,2,16,8,11 >> $O/r2h.jsonl 2>>$O/r2h_err.txt
cat $O/r2h.jsonl | cut -c1-30,80-290
