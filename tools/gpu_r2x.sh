#!/bin/bash
# round 2, GPU call X: PDW median on |y|^2 with the lean binning loop and a one-wave grid: tests + per-stage GPU times
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2x_*
timeout 900 python -m pytest tests/test_gpu_pdw.py tests/test_sharding.py -x -q -m gpu > $O/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2x_status.txt
timeout 300 python tools/exp/pdw_trace.py 8 >> $O/r2x_trace.jsonl 2>> $O/r2x.err
for c in 64 32 16 128; do
  echo "chunk_rows=$c" >> $O/r2x_trace.jsonl
  CHZ_PDW_CHUNK_ROWS=$c CHZ_PDW_TRACE=1 timeout 300 python tools/exp/pdw_trace.py 2 2>&1 | grep "pdw gpu" | tail -n 2 >> $O/r2x_trace.jsonl
done
tail -n 3 $O/r2x_pytest.log; cat $O/r2x_trace.jsonl; cat $O/r2x_status.txt; tail -n 3 $O/r2x.err
