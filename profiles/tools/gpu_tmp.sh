#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detection.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/g_tests.log 2>&1
echo "rc $?" >> gpurun_out/g_tests.log
tail -3 gpurun_out/g_tests.log
for v in "" r1 r2 m3 l8; do
  echo "== variant $v"
  if [ -n "$v" ]; then export TF_LIB_PATH=$PWD/profiles/tools/_var/libtf_$v.so; fi
  python profiles/tools/gather_time.py 2>&1 | grep "nans=True"
done
unset TF_LIB_PATH
echo "== general"
TF_GATHER_GENERAL=1 TF_SOBEL_GENERAL=1 python profiles/tools/gather_time.py 2>&1 | grep "nans=True"
