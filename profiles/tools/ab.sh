#!/bin/bash
# usage: profiles/tools/ab.sh "ENV1=a ENV2=b" "ENV1=c" ...  -> runs bench (96 frames, no e2e/cpu) per env setting
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  echo "=== $envs" >> gpurun_out/ab.log
  env $envs python bench.py --frames 96 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | python profiles/tools/ab_parse.py >> gpurun_out/ab.log
  i=$((i+1))
done
cat gpurun_out/ab.log
