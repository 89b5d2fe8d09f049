#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/l_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/l_smoke.log
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/l_tests.log 2>&1
echo "rc $?" >> gpurun_out/l_tests.log
tail -3 gpurun_out/l_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err
echo "bench rc $?"; tail -2 gpurun_out/bench_r2_final2.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final2.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'whole',r['whole_step']['frac'],'dom',r['frac'],'launches',d['gpu_launches'])
print('e2e',{k:v for k,v in d['e2e'].items() if k!='note'})
tot=0
for k,v in r['per_class'].items(): print('  %-16s %8.2f ms %s'%(k,v['ms_per_step'],v['GBps'])); tot+=v['ms_per_step']
print('classes',tot, d['clocks'])
PY
