#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k "full_disk_stencils" > gpurun_out/s22_tests.log 2>&1
echo "rc $?" >> gpurun_out/s22_tests.log
tail -6 gpurun_out/s22_tests.log
python -c "
import __graft_entry__ as g
g.smoke()" > gpurun_out/s22_smoke.log 2>&1
echo "smoke rc $?"; tail -2 gpurun_out/s22_smoke.log
free -g | head -2
