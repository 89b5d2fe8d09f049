#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/s11_tests.log 2>&1
echo "rc $?" >> gpurun_out/s11_tests.log
tail -4 gpurun_out/s11_tests.log
bash scratch/ab.sh "TF_X=1" > gpurun_out/s11_ab_stdout.log 2>&1
grep -E "===|fps|pyramid|fb_iter" gpurun_out/ab.log
python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s11_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:blur -s 6 -c 6 --csv --log-file gpurun_out/s11_pyr_launches.csv python bench.py --frames 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s11_ncu.log 2>&1
grep -E "blur" gpurun_out/s11_pyr_launches.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $(NF)}' | head -30
