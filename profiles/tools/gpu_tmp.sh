#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/n_tests.log 2>&1
echo "rc $?" >> gpurun_out/n_tests.log; tail -2 gpurun_out/n_tests.log
for i in 1 2; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_final3_$i.json 2> gpurun_out/bench_r2_final3_$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_final3_$i.json').read().strip().splitlines()[-1])
r=d['roofline']
tot=sum(v['ms_per_step'] for v in r['per_class'].values())
print('run $i value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'classes',round(tot,1),'whole',round(r['whole_step']['frac'],4),'dom',round(r['frac'],4),'e2e',round(d['e2e']['value'],1),round(d['e2e']['value_diff_sobel_convolve_order'],1), d['clocks']['sm_mhz'])
PY
done
