#!/bin/bash
# round 2, GPU call Y: full GPU suite + bench.py after the PDW / small-call changes
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/r2y4_*
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2y4_pytest_all.log 2>&1
echo "pytest-all rc=$?" >> $O/r2y4_status.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2y4_bench_main.json 2> $O/r2y4_bench_main.err
echo "bench rc=$?" >> $O/r2y4_status.txt
timeout 300 python tools/exp/small_file_latency.py >> $O/r2y4_small.jsonl 2>> $O/r2y.err
tail -n 3 $O/r2y4_pytest_all.log; cat $O/r2y4_status.txt; cat $O/r2y4_small.jsonl; tail -n 3 $O/r2y4_bench_main.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2y4_bench_main.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline'].get('same_traffic_noop'))
for o in d['others']: print(o['config'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in o.items() if k not in ('config','note')})
print(d['e2e']['value'], d['e2e_pdw']['value'])
PY
