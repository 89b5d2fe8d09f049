#!/bin/bash
rm -f gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/d_tests.log 2>&1
echo "rc $?" >> gpurun_out/d_tests.log; tail -4 gpurun_out/d_tests.log
bash profiles/tools/ab.sh "TF_X=1" "TF_PYR_NO_EXACT=1" > gpurun_out/d_ab_stdout.log 2>&1
grep -E "===|fps|pyramid" gpurun_out/ab.log
