#!/bin/bash
mkdir -p gpurun_out
for c in c1 c3 c4 c5; do
python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-detection --e2e-frames 24 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err
echo "$c rc $?"; tail -1 gpurun_out/bench_r2_$c.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_$c.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$c', 'frames/s %.1f'%d['value'],'ms %.1f'%d['ms_per_step'],'whole %.3f'%r['whole_step']['frac'],'dom %.3f'%r['frac'], 'e2e', round(d['e2e']['value'],1) if d['e2e'] else None, d['config']['workload'][:70])
PY
done
