#!/bin/bash
# round 2, GPU call V: PDW extractor as a CUDA graph: tests, per-file latency with and without, multi-file bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2v_*
timeout 900 python -m pytest tests/test_gpu_pdw.py tests/test_sharding.py -x -q -m gpu > $O/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2v_status.txt
for g in 1 0; do
  CHZ_PDW_GRAPH=$g timeout 300 python tools/exp/pdw_trace.py 8 >> $O/r2v_trace.jsonl 2>> $O/r2v.err
done
CHZ_PDW_TRACE=1 timeout 300 python tools/exp/pdw_trace.py 2 2>&1 | grep "pdw gpu" | tail -n 3 >> $O/r2v_trace.jsonl
tail -n 3 $O/r2v_pytest.log; cat $O/r2v_trace.jsonl; cat $O/r2v_status.txt; tail -n 3 $O/r2v.err
timeout 200 python tools/exp/small_file_latency.py >> $O/r2v_small.jsonl 2>> $O/r2v.err
M=64 timeout 200 python tools/exp/small_file_latency.py >> $O/r2v_small.jsonl 2>> $O/r2v.err
M=1024 timeout 200 python tools/exp/small_file_latency.py >> $O/r2v_small.jsonl 2>> $O/r2v.err
timeout 900 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu > $O/r2v_pytest_chan.log 2>&1
echo "pytest-chan rc=$?" >> $O/r2v_status.txt
cat $O/r2v_small.jsonl; tail -n 2 $O/r2v_pytest_chan.log; cat $O/r2v_status.txt
