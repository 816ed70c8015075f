#!/bin/bash
# round 2, GPU call R: A/B of ws ring kernel build variants (libchannelizer-<name>.so via CHZ_LIB_PATH)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2r_*
export CHZ_RING_VARIANT=2
for v in ${VARS:-"" -regs -r128 -r144 -r160}; do
  echo "variant=$v" >> $O/r2r_bench.jsonl
  CHZ_LIB_PATH=$PWD/sdr_channelizer_b200/libchannelizer$v.so timeout 300 python tools/exp/bench_paths.py ${SPECS:-1024,1,16,16,0 1024,2,16,16,0 1024,1,12,12,0 1024,1,8,16,0} >> $O/r2r_bench.jsonl 2>> $O/r2r_bench.err
done
cat $O/r2r_bench.jsonl; tail -n 5 $O/r2r_bench.err
