#!/bin/bash
mkdir -p gpurun_out profiles
rm -f gpurun_out/ab.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s21_gpu_tests.log 2>&1
echo "rc $?" >> gpurun_out/s21_gpu_tests.log
tail -4 gpurun_out/s21_gpu_tests.log
bash scratch/ab.sh "TF_X=1" > gpurun_out/s21_ab_stdout.log 2>&1
grep -E "===|fps|sl_gather|fb_iter" gpurun_out/ab.log
python bench.py --config c1 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s21_c1_plain.log 2>&1
echo "c1 plain rc $?"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck_c1.log python bench.py --config c1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-detection > gpurun_out/s21_c1_memcheck.out 2>&1
echo "memcheck rc $?"
tail -5 gpurun_out/r2_sanitizer_memcheck_c1.log
