#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
V=$PWD/scratch/_var
bash scratch/ab.sh "TF_TMEM=1" "TF_TMEM=1 TF_LIB_PATH=$V/libtf_abl4.so" "TF_TMEM=0 TF_LIB_PATH=$V/libtf_abl4.so" > gpurun_out/s6_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter" gpurun_out/ab.log
for m in 1 2; do
TF_TMA=$m python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s6_plain.log 2>&1 && \
TF_TMA=$m ncu --set full --clock-control none --import-source on -k regex:fb_iter_tma -s 112 -c 1 -o gpurun_out/prof_fb_tma${m}_r2 -f python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s6_ncu.log 2>&1
done
tail -2 gpurun_out/s6_ncu.log
