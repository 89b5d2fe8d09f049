#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s10_tests.log 2>&1
echo "rc $?" >> gpurun_out/s10_tests.log
tail -4 gpurun_out/s10_tests.log
bash scratch/ab.sh "TF_X=1" "TF_PYR_TWO_PASS=1" > gpurun_out/s10_ab_stdout.log 2>&1
grep -E "===|fps|pyramid|fb_iter" gpurun_out/ab.log
python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s10_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.avg.per_cycle_elapsed --clock-control none -k regex:blur -s 6 -c 6 --csv --log-file gpurun_out/s10_pyr_launches.csv python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s10_ncu.log 2>&1
grep -E "blur" gpurun_out/s10_pyr_launches.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | head -30
