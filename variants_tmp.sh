python tools/bench_configs.py --only cfg3,cfg4 --scale 0.25 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d.get('config'), d.get('error') or (round(d['ms_per_pass'],3), round(d['MS_per_s']), round(d['frac_of_measured_hbm'],3), d['launches_per_pass']))
"
