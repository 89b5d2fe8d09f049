#!/bin/bash
rm -f gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/b_tests.log 2>&1
echo "rc $?" >> gpurun_out/b_tests.log
tail -4 gpurun_out/b_tests.log
bash profiles/tools/ab.sh "TF_X=1" "TF_TMA=0" > gpurun_out/b_ab_stdout.log 2>&1
grep -E "===|fps|fb_iter" gpurun_out/ab.log
