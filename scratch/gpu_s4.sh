#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
for m in 1 2; do
TF_TMA=$m timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/s4_parity_tma$m.log 2>&1
echo "parity rc $?" >> gpurun_out/s4_parity_tma$m.log
tail -4 gpurun_out/s4_parity_tma$m.log
done
bash scratch/ab.sh "TF_TMEM=1" "TF_TMA=1" "TF_TMA=2" "TF_TMEM=1" "TF_TMA=1" "TF_TMA=2" > gpurun_out/s4_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter" gpurun_out/ab.log
tail -3 gpurun_out/s4_ab_stdout.log
