#!/bin/bash
# round 2, GPU call T: split path (k_fir + k_fft_rows_big) on configs[3] geometry processed in row chunks: DRAM bytes
# per kernel with ncu's cache flush between kernels switched off (does the FIR output stay in L2 for the FFT kernel?)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2t_ncu*
for c in 0 2048 1024 512 256; do
  STEPS=1 timeout 600 ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file $O/r2t_ncu_$c.csv python tools/exp/chunked_split.py 4096 16 $c 140000000 > /dev/null 2>> $O/r2t.err
done
tail -n 3 $O/r2t.err; wc -l $O/r2t_ncu_*.csv
