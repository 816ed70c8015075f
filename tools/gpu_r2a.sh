#!/bin/bash
# round 2, GPU call A: first run of the ring kernel (M = 1024): parity tests, path A/B, sanitizer
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu -k "1024" > $O/r2a_pytest_1024.log 2>&1
echo "pytest-1024 rc=$?" >> $O/r2a_status.txt
for u in 0 1 2; do
  CHZ_RING_UNPACK=$u timeout 300 python tools/exp/bench_paths.py 1024,2,16,16,11 1024,1,16,16,11 1024,1,12,12,11 1024,2,16,8,11 >> $O/r2a_bench.jsonl 2>> $O/r2a_bench.err
done
timeout 300 python tools/exp/bench_paths.py 1024,2,16,16,2 1024,1,16,16,2 64,1,16,12,0,614400000 >> $O/r2a_bench.jsonl 2>> $O/r2a_bench.err
echo "bench rc=$?" >> $O/r2a_status.txt
timeout 600 compute-sanitizer --tool memcheck python tools/exp/bench_paths.py 1024,2,16,16,11,4000000 1024,1,12,8,11,4000000 > $O/r2a_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/r2a_status.txt
timeout 600 compute-sanitizer --tool racecheck python tools/exp/bench_paths.py 1024,2,16,16,11,2000000 > $O/r2a_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/r2a_status.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2a_pytest_all.log 2>&1
echo "pytest-all rc=$?" >> $O/r2a_status.txt
tail -3 $O/r2a_pytest_1024.log; cat $O/r2a_bench.jsonl; tail -5 $O/r2a_memcheck.log; tail -5 $O/r2a_racecheck.log; tail -3 $O/r2a_pytest_all.log; cat $O/r2a_status.txt
