#!/bin/bash
rm -f gpurun_out/ab.log
V=$PWD/profiles/tools/_var
TF_TMA=4 TF_LIB_PATH=$V/libtf_occ5.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not kernels_agree" > gpurun_out/c_tests.log 2>&1
echo "rc $?" >> gpurun_out/c_tests.log; tail -3 gpurun_out/c_tests.log
bash profiles/tools/ab.sh "TF_X=1" "TF_TMA=4 TF_LIB_PATH=$V/libtf_occ5.so" "TF_TMA=3 TF_LIB_PATH=$V/libtf_occ5.so" "TF_TMA=4" > gpurun_out/c_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter" gpurun_out/ab.log
