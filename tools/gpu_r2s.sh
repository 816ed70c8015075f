#!/bin/bash
# round 2, GPU call S: role-split ring kernel as the default M = 1024 path: full GPU suite, A/B lines, bench.py
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2s_*
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2s_pytest_all.log 2>&1
echo "pytest-all rc=$?" >> $O/r2s_status.txt
timeout 300 python tools/exp/bench_paths.py 1024,1,16,16,0 1024,2,16,16,0 1024,1,12,12,0 1024,2,16,8,0 1024,1,8,16,0 1024,2,12,12,0 1024,1,16,8,0 >> $O/r2s_bench.jsonl 2>> $O/r2s_bench.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2s_bench_main.json 2> $O/r2s_bench_main.err
echo "bench rc=$?" >> $O/r2s_status.txt
tail -n 3 $O/r2s_pytest_all.log; cat $O/r2s_bench.jsonl; cat $O/r2s_status.txt; tail -n 3 $O/r2s_bench_main.err; head -c 3000 $O/r2s_bench_main.json
