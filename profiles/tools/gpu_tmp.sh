#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fb_iter|sl_gather|sl_lean|sobel_lin|pyr_|blur|polyexp|flow_upsample|resize_tables|pair_|minmax|finalise" -c 400 --csv --log-file gpurun_out/launches_r2_final.csv $B > gpurun_out/ncu_launches_final.log 2>&1
tail -1 gpurun_out/ncu_launches_final.log | cut -c1-100
