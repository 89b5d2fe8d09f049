#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection"
$B > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fb_iter_v3 -s 50 -c 2 -o gpurun_out/r2_prof_fb_v3_up -f $B > gpurun_out/ncu_fb_up.log 2>&1
tail -2 gpurun_out/ncu_fb_up.log | cut -c1-200
