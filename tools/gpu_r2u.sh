#!/bin/bash
# round 2, GPU call U: split path with rotated rows (k_fir ROT): parity tests + configs[3] full size + small-size A/B
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2u_*
timeout 900 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu -k "1024 or 2048 or 4096 or split or shard or stream" > $O/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2u_status.txt
timeout 300 python tools/exp/bench_paths.py 4096,1,16,16,0 4096,2,16,16,0 2048,1,16,16,0 4096,1,12,12,0 1024,1,16,16,2 4096,1,16,16,0,3686400000 >> $O/r2u_bench.jsonl 2>> $O/r2u_bench.err
tail -n 3 $O/r2u_pytest.log; cat $O/r2u_bench.jsonl; cat $O/r2u_status.txt; tail -n 3 $O/r2u_bench.err
