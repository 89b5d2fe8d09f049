#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu -k "multi or sharded or finalise" > gpurun_out/s20_multi_tests.log 2>&1
echo "rc $?" >> gpurun_out/s20_multi_tests.log
tail -15 gpurun_out/s20_multi_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err
echo "bench rc $?"
tail -3 gpurun_out/bench_r2_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_2gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'] if d['e2e'] else None,'parity',d.get('sharded_parity'),'strong',d.get('strong'))
print(d['config']['host_pinning'])
PY
