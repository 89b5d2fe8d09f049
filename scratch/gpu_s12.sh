#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
echo "rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_a.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'whole',r['whole_step']['frac'],'dom',r['frac'])
for k,v in r['per_class'].items(): print('  %-16s %8.2f ms %s'%(k,v['ms_per_step'],v['GBps']))
print(d['cpu_baseline']); print(d['clocks']); print(d['config']['host_pinning'])
PY
tail -5 gpurun_out/bench_r2_a.err
