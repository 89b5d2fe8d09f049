#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detection.py -x -q -m gpu > gpurun_out/j_tests.log 2>&1
echo "rc $?" >> gpurun_out/j_tests.log
tail -25 gpurun_out/j_tests.log
