#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
V=$PWD/scratch/_var
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/s15_tests.log 2>&1
echo "rc $?" >> gpurun_out/s15_tests.log
tail -3 gpurun_out/s15_tests.log
bash scratch/ab.sh "TF_LIB_PATH=$V/libtf_v3base.so" "TF_X=1" "TF_X=1" > gpurun_out/s15_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter|pyramid|polyexp|upsample" gpurun_out/ab.log
