#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/s13_tests.log 2>&1
echo "rc $?" >> gpurun_out/s13_tests.log
tail -3 gpurun_out/s13_tests.log
for l in 1 2 1 2; do
echo "=== lanes $l"
TF_FLOW_LANES=$l python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-detection 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fps %.1f ms %.1f'%(d['value'],d['ms_per_step'])); print({k:round(v['ms_per_step'],1) for k,v in d['roofline']['per_class'].items()})
    else: print(l[:200])
"
done
