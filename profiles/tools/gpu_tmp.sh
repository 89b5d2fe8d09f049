#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fb_iter|sl_gather|sl_lean|sobel_lin|pyr_|blur|polyexp|flow_upsample|pair_|minmax|finalise" -c 400 --csv --log-file gpurun_out/launches_r2_final.csv $B > gpurun_out/ncu_launches_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sobel_lin|sl_lean|pyr_half|pyr_row" -c 9 -o gpurun_out/r2_prof_small_final -f $B > gpurun_out/ncu_small_final.log 2>&1
tail -2 gpurun_out/ncu_small_final.log | cut -c1-200
