#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/s_smoke.log 2>&1; echo "smoke rc $?"
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/s_tests.log 2>&1
echo "rc $?" >> gpurun_out/s_tests.log
tail -3 gpurun_out/s_tests.log
