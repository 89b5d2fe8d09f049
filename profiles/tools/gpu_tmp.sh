#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detection.py -x -q -m gpu > gpurun_out/u_tests.log 2>&1
echo "rc $?" >> gpurun_out/u_tests.log; tail -2 gpurun_out/u_tests.log
python profiles/tools/gather_time.py 2>&1 | grep "nans=True"
