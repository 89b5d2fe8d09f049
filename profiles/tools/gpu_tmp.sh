#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/g_tests.log 2>&1
echo "rc $?" >> gpurun_out/g_tests.log
tail -5 gpurun_out/g_tests.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err
echo "bench rc $?"; tail -3 gpurun_out/bench_r2_c.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_c.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_step']['frac'])
print('e2e',{k:v for k,v in d['e2e'].items() if k!='note'})
for k,v in d['roofline']['per_class'].items(): print("  %-16s %8.2f ms %6.0f GB/s"%(k,v['ms_per_step'],v['GBps']))
print(d['clocks'], d.get('detection'))
PY
