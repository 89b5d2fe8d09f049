#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
V=$PWD/scratch/_var
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/s3_stages.log 2>&1
echo "stages rc $?" >> gpurun_out/s3_stages.log
tail -15 gpurun_out/s3_stages.log
bash scratch/ab.sh "TF_TMEM=1" "TF_TMEM=1 TF_PYR_TWO_PASS=1" "TF_TMEM=1 TF_LIB_PATH=$V/libtf_abl1.so" "TF_TMEM=1 TF_LIB_PATH=$V/libtf_abl2.so" "TF_TMEM=1 TF_LIB_PATH=$V/libtf_abl3.so" "TF_TMEM=0 TF_LIB_PATH=$V/libtf_abl2.so" > gpurun_out/s3_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter|pyramid" gpurun_out/ab.log
tail -5 gpurun_out/s3_ab_stdout.log
