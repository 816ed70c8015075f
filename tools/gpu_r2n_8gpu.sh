#!/bin/bash
# round 2, 8-GPU call: end-to-end scaling with NUMA binding, configs[4] replicas, sharded PDW with typed gathers
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-8}
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; numactl -H 2>/dev/null | head -20) > $O/r2n_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r2n_bench_${N}gpu.json 2> $O/r2n_bench_${N}gpu.err
echo "bench rc=$?" > $O/r2n_status.txt
FILES_PER_RANK=64 WORKERS=4 $TR tools/bench_cfg5_dist.py > $O/r2n_cfg5_${N}gpu.json 2>> $O/r2n_bench_${N}gpu.err
echo "cfg5 rc=$?" >> $O/r2n_status.txt
FILES_PER_RANK=64 WORKERS=1 $TR tools/bench_cfg5_dist.py > $O/r2n_cfg5_${N}gpu_1worker.json 2>> $O/r2n_bench_${N}gpu.err
$TR tools/run_sharded_pdw.py > $O/r2n_sharded_pdw_${N}gpu.json 2>> $O/r2n_bench_${N}gpu.err
echo "sharded rc=$?" >> $O/r2n_status.txt
cat $O/r2n_status.txt; tail -3 $O/r2n_bench_${N}gpu.err; cat $O/r2n_cfg5_${N}gpu.json $O/r2n_cfg5_${N}gpu_1worker.json $O/r2n_sharded_pdw_${N}gpu.json
