#!/bin/bash
# session 1: TMEM-ring A/B
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
TF_TMEM=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/s1_parity_tmem.log 2>&1
echo "parity rc $?" >> gpurun_out/s1_parity_tmem.log
bash scratch/ab.sh "TF_TMEM=0" "TF_TMEM=1" "TF_TMEM=0" "TF_TMEM=1" > /dev/null 2>&1
TF_TMEM=1 python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s1_plain.log 2>&1 && \
TF_TMEM=1 ncu --set full --clock-control none --import-source on -k regex:fb_iter_strip -s 112 -c 1 -o gpurun_out/prof_fb_tmem_r2 -f python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s1_ncu.log 2>&1
tail -3 gpurun_out/s1_parity_tmem.log
cat gpurun_out/ab.log
