#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
SPECS="64,1,16,12,0,614400000 64,1,12,12,0,614400000 32,1,16,12,0 128,1,16,12,0 128,1,12,12,0 256,1,16,16,0 512,1,16,12,0 8,1,8,8,0 16,1,16,12,0 56,1,12,16,0 560,1,12,16,0 4096,1,16,12,0 2048,1,16,12,0 1024,2,16,16,0"
rm -f $O/r2o_cmul_ab.jsonl
for lib in libchannelizer.so libchannelizer-pcmul.so; do
  CHZ_LIB_PATH=$PWD/sdr_channelizer_b200/$lib python tools/exp/bench_paths.py $SPECS | sed "s/^{/{\"lib\": \"$lib\", /" >> $O/r2o_cmul_ab.jsonl
done
CHZ_LIB_PATH=$PWD/sdr_channelizer_b200/libchannelizer-pcmul.so python -m pytest tests/test_gpu_channelizer.py -x -q -m gpu -k "matches_oracle or cufft or random_taps" 2>&1 | tail -2
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2o_cmul_ab.jsonl')]
a={ (r['M'],r['oversample'],r['P'],r['bits']):r for r in rows if r['lib']=='libchannelizer.so'}
b={ (r['M'],r['oversample'],r['P'],r['bits']):r for r in rows if r['lib']!='libchannelizer.so'}
for k in a:
    print(k, 'scalar cmul %.4f ms %.3f | packed %.4f ms %.3f | %+.1f%%' % (a[k]['ms'], a[k]['frac_of_measured_hbm'], b[k]['ms'], b[k]['frac_of_measured_hbm'], 100*(a[k]['ms']/b[k]['ms']-1)))
PY
