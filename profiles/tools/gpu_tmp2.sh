#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/t_multi_tests.log 2>&1
echo "rc $?" >> gpurun_out/t_multi_tests.log; tail -3 gpurun_out/t_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r2_2gpu_final2.json 2> gpurun_out/bench_r2_2gpu_final2.err
echo "bench rc $?"; tail -2 gpurun_out/bench_r2_2gpu_final2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_2gpu_final2.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',{k:v for k,v in d['e2e'].items() if k!='note'},'parity',d.get('sharded_parity'),'strong',{k:v for k,v in d.get('strong',{}).items() if k!='note'})
PY
