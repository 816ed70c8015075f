run() { # sr lag blocks
  CHZ_PIPE_SPAN_ROWS=$1 CHZ_PIPE_LAG=$2 CHZ_PIPE_BLOCKS=$3 CHZ_BENCH_PATH=6 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_chan_pipe -s 1 -c 1 --csv python tools/bench_configs.py --only cfg4 --scale 0.1 --steps 1 2>&1 | grep -E '"dram__bytes|gpu__time' | awk -F'","' -v tag="sr=$1 lag=$2 blocks=$3" '{printf "%s %s %s %s\n", tag, $(NF-2), $(NF-1), $NF}'
}
run 64 0 0
run 64 4 0
run 64 2 148
run 64 6 148
run 32 4 148
run 32 8 296
run 16 8 296
run 16 4 148
