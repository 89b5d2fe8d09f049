#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
V=$PWD/profiles/tools/_var/libtf_up2.so
TF_LIB_PATH=$V timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "iteration or flow_small or three_levels or odd_sizes" > gpurun_out/r_tests.log 2>&1
echo "rc $?" >> gpurun_out/r_tests.log; tail -2 gpurun_out/r_tests.log
bash profiles/tools/ab.sh "TF_X=1" "TF_LIB_PATH=$V" "TF_X=1" "TF_LIB_PATH=$V" > /dev/null 2>&1
grep -E "===|fps|fb_iter" gpurun_out/ab.log
