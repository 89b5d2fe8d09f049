#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_r2_8gpu_final.json 2> gpurun_out/bench_r2_8gpu_final.err
echo "bench rc $?"; tail -3 gpurun_out/bench_r2_8gpu_final.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_8gpu_final.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',{k:v for k,v in d['e2e'].items() if k!='note'},'parity',d.get('sharded_parity'),'strong',d.get('strong'))
print(d['config']['host_pinning'], d['clocks'])
PY
