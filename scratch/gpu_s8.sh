#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
TF_TMA=3 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s8_parity.log 2>&1
echo "parity rc $?" >> gpurun_out/s8_parity.log
tail -4 gpurun_out/s8_parity.log
bash scratch/ab.sh "TF_TMA=2" "TF_TMA=3" "TF_TMA=4" "TF_TMA=2" "TF_TMA=3" > gpurun_out/s8_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter" gpurun_out/ab.log
TF_TMA=3 python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s8_plain.log 2>&1 && \
TF_TMA=3 ncu --set full --clock-control none --import-source on -k regex:fb_iter_v3 -s 112 -c 1 -o gpurun_out/prof_fb_v3_r2 -f python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s8_ncu.log 2>&1
tail -2 gpurun_out/s8_ncu.log
