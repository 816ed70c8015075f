#!/bin/bash
# round 2, second 8-GPU call: configs[4] replicas and the sharded PDW extraction with the reworked PDW kernels
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
mkdir -p $O
rm -f $O/r2z_*
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
FILES_PER_RANK=64 WORKERS=4 timeout 600 $TR tools/bench_cfg5_dist.py > $O/r2z_cfg5_${N}gpu.json 2> $O/r2z_${N}gpu.err
echo "cfg5 rc=$?" >> $O/r2z_status.txt
FILES_PER_RANK=64 WORKERS=1 timeout 600 $TR tools/bench_cfg5_dist.py > $O/r2z_cfg5_${N}gpu_1worker.json 2>> $O/r2z_${N}gpu.err
echo "cfg5-1w rc=$?" >> $O/r2z_status.txt
timeout 600 $TR tools/run_sharded_pdw.py > $O/r2z_sharded_pdw_${N}gpu.json 2>> $O/r2z_${N}gpu.err
echo "sharded rc=$?" >> $O/r2z_status.txt
cat $O/r2z_status.txt; tail -n 3 $O/r2z_${N}gpu.err; cat $O/r2z_cfg5_${N}gpu.json $O/r2z_cfg5_${N}gpu_1worker.json $O/r2z_sharded_pdw_${N}gpu.json
