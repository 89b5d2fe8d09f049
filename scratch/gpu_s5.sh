#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
TF_PK=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s5_parity_pk.log 2>&1
echo "parity rc $?" >> gpurun_out/s5_parity_pk.log
tail -4 gpurun_out/s5_parity_pk.log
bash scratch/ab.sh "TF_TMEM=1" "TF_PK=1" "TF_TMEM=1" "TF_PK=1" > gpurun_out/s5_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter" gpurun_out/ab.log
TF_PK=1 python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s5_plain.log 2>&1 && \
TF_PK=1 ncu --set full --clock-control none --import-source on -k regex:fb_iter_pk -s 112 -c 1 -o gpurun_out/prof_fb_pk_r2 -f python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s5_ncu.log 2>&1
tail -2 gpurun_out/s5_ncu.log
