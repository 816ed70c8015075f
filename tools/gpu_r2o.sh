#!/bin/bash
# round 2, GPU call O: warp-specialised ring kernel (CHZ_RING_VARIANT=2): parity + A/B against the default ring kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2o_*
CHZ_RING_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu -k "1024" > $O/r2o_pytest_1024.log 2>&1
echo "pytest-1024 rc=$?" >> $O/r2o_status.txt
for v in 0 2; do
  CHZ_RING_VARIANT=$v timeout 300 python tools/exp/bench_paths.py 1024,1,16,16,0 1024,2,16,16,0 1024,1,12,12,0 1024,2,16,8,0 1024,1,8,16,0 >> $O/r2o_bench.jsonl 2>> $O/r2o_bench.err
done
tail -n 3 $O/r2o_pytest_1024.log; cat $O/r2o_bench.jsonl; cat $O/r2o_status.txt; tail -n 5 $O/r2o_bench.err
