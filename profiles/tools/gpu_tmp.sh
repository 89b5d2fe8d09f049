#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/o_tests.log 2>&1
echo "rc $?" >> gpurun_out/o_tests.log; tail -4 gpurun_out/o_tests.log
