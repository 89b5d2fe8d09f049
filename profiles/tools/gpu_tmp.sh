#!/bin/bash
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/e_tests.log 2>&1
echo "rc $?" >> gpurun_out/e_tests.log; tail -5 gpurun_out/e_tests.log
python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/e_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fb_iter|sl_gather|polyexp|blur|flow_upsample|flow_finalise|pair_|minmax' -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/e_ncu.log 2>&1
tail -1 gpurun_out/e_ncu.log
