#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
V=$PWD/scratch/_var
bash scratch/ab.sh "TF_X=1" "TF_LIB_PATH=$V/libtf_gmb3.so" "TF_LIB_PATH=$V/libtf_gmb5.so" "TF_LIB_PATH=$V/libtf_gmb6.so" "TF_PAIR_BATCH_MPX=512" "TF_PAIR_BATCH_MPX=128" > gpurun_out/s18_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter|sl_gather" gpurun_out/ab.log
