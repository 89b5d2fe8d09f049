#!/bin/bash
mkdir -p gpurun_out
for c in c1 c3 c4 c5; do
python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-detection > gpurun_out/bench_r2_final_$c.json 2> gpurun_out/bench_r2_final_$c.err
echo "$c rc $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_final_$c.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$c', d['config']['workload'][:60], 'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'whole',round(r['whole_step']['frac'],3),'dom',round(r['frac'],3),'e2e',round(d['e2e']['value'],1) if d.get('e2e') else None)
PY
done
