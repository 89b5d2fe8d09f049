#!/bin/bash
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/a_gpu_tests.log 2>&1
echo "rc $?" >> gpurun_out/a_gpu_tests.log
tail -4 gpurun_out/a_gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_1gpu_b.json 2> gpurun_out/bench_r2_1gpu_b.err
echo "bench rc $?"; tail -2 gpurun_out/bench_r2_1gpu_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_1gpu_b.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['frames_per_step'],'whole',r['whole_step']['frac'],'dom',r['frac'])
for k,v in r['per_class'].items(): print('  %-16s %8.2f ms %s'%(k,v['ms_per_step'],v['GBps']))
print(d['clocks'])
PY
