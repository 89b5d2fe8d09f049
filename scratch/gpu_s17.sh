#!/bin/bash
mkdir -p gpurun_out
python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s17_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fb_iter_v3 -s 112 -c 1 -o gpurun_out/prof_fb_v3b_r2 -f python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s17_ncu.log 2>&1
tail -2 gpurun_out/s17_ncu.log
