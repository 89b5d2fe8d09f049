#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s19_gpu_tests.log 2>&1
echo "rc $?" >> gpurun_out/s19_gpu_tests.log
tail -4 gpurun_out/s19_gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2_1gpu.json 2> gpurun_out/bench_r2_1gpu.err
echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'whole',r['whole_step']['frac'],'dom',r['frac'])
for k,v in r['per_class'].items(): print('  %-16s %8.2f ms %s'%(k,v['ms_per_step'],v['GBps']))
print(d['clocks'])
PY
python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s19_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s19_ncu.log 2>&1
tail -1 gpurun_out/s19_ncu.log
