#!/bin/bash
# round 2, GPU call AA: 16-byte paired stores in the last FFT pass (libchannelizer-pair.so) against the default build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2aa_*
CHZ_LIB_PATH=$PWD/sdr_channelizer_b200/libchannelizer-pair.so timeout 900 python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu > $O/r2aa_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2aa_status.txt
for v in "" -pair "" -pair; do
  echo "variant=$v" >> $O/r2aa_bench.jsonl
  CHZ_LIB_PATH=$PWD/sdr_channelizer_b200/libchannelizer$v.so timeout 300 python tools/exp/bench_paths.py 64,1,16,12,0,614400000 64,1,12,12,0,614400000 32,1,16,12,0 128,1,12,12,0 256,1,16,16,0 512,1,16,16,0 64,2,16,12,0 4096,1,16,12,0 >> $O/r2aa_bench.jsonl 2>> $O/r2aa.err
done
tail -n 2 $O/r2aa_pytest.log; cat $O/r2aa_status.txt; cat $O/r2aa_bench.jsonl; tail -n 3 $O/r2aa.err
