#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/q_tests.log 2>&1
echo "rc $?" >> gpurun_out/q_tests.log
tail -3 gpurun_out/q_tests.log
