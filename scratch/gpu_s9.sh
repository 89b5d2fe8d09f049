#!/bin/bash
mkdir -p gpurun_out
python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s9_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blur -s 6 -c 6 -o gpurun_out/prof_pyr_r2 -f python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s9_ncu.log 2>&1
tail -2 gpurun_out/s9_ncu.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s9_gpu_tests.log 2>&1
echo "rc $?" >> gpurun_out/s9_gpu_tests.log
tail -5 gpurun_out/s9_gpu_tests.log
