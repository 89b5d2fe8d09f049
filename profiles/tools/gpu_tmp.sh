#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_final5.json 2> gpurun_out/bench_r2_final5.err
echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final5.json').read().strip().splitlines()[-1])
r=d['roofline']
tot=sum(v['ms_per_step'] for v in r['per_class'].values())
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'classes',round(tot,1),'whole',round(r['whole_step']['frac'],4),'dom',round(r['frac'],4),'launches',d['gpu_launches'])
print('e2e',{k:v for k,v in d['e2e'].items() if k!='note'})
for k,v in r['per_class'].items(): print('  %-16s %8.2f ms %s'%(k,v['ms_per_step'],round(v['GBps'])))
print(d['clocks'])
PY
