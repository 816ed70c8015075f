#!/bin/bash
# round 2, GPU call AC: ring kernel with 16 FFT warps (CHZ_RING_VARIANT=3) against the default 8; every command under a
# short timeout (a wrong setmaxnreg budget deadlocked the first attempt)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2ac_*
CHZ_RING_VARIANT=3 STEPS=2 timeout 40 python tools/exp/bench_paths.py 1024,1,16,16,0,28000000 >> $O/r2ac_bench.jsonl 2>> $O/r2ac.err
echo "first rc=$?" >> $O/r2ac_status.txt
if grep -q GS_per_s $O/r2ac_bench.jsonl; then
  CHZ_RING_VARIANT=3 timeout 120 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu -k "1024" > $O/r2ac_pytest.log 2>&1
  echo "pytest rc=$?" >> $O/r2ac_status.txt
  for v in 3 0; do
    CHZ_RING_VARIANT=$v timeout 60 python tools/exp/bench_paths.py 1024,1,16,16,0 1024,2,16,16,0 1024,1,12,12,0 1024,2,12,12,0 1024,1,8,16,0 >> $O/r2ac_bench.jsonl 2>> $O/r2ac.err
  done
fi
tail -n 2 $O/r2ac_pytest.log; cat $O/r2ac_status.txt; cat $O/r2ac_bench.jsonl; tail -n 3 $O/r2ac.err
