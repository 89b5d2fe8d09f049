#!/bin/bash
# session 2: deep-lookahead schedule A/B (x TMEM ring)
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
V=$PWD/scratch/_var
TF_TMEM=1 TF_LIB_PATH=$V/libtf_deep2.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s2_parity.log 2>&1
echo "parity rc $?" >> gpurun_out/s2_parity.log
bash scratch/ab.sh "TF_TMEM=1" "TF_TMEM=1 TF_LIB_PATH=$V/libtf_deep2.so" "TF_TMEM=0 TF_LIB_PATH=$V/libtf_deep2.so" "TF_TMEM=1 TF_LIB_PATH=$V/libtf_deep4.so" "TF_TMEM=1" "TF_TMEM=1 TF_LIB_PATH=$V/libtf_deep2.so" > /dev/null 2>&1
TF_TMEM=1 TF_LIB_PATH=$V/libtf_deep2.so python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s2_plain.log 2>&1 && \
TF_TMEM=1 TF_LIB_PATH=$V/libtf_deep2.so ncu --set full --clock-control none --import-source on -k regex:fb_iter_strip -s 112 -c 1 -o gpurun_out/prof_fb_deep2_r2 -f python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s2_ncu.log 2>&1
tail -3 gpurun_out/s2_parity.log
cat gpurun_out/ab.log
