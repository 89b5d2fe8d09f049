#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_last.json 2> gpurun_out/bench_r2_last.err
echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_last.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'whole',round(r['whole_step']['frac'],4),'dom',round(r['frac'],4),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'], d['clocks']['sm_mhz'])
PY
