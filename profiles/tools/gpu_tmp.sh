#!/bin/bash
for v in "" b4m8 b4m9 b4m10 b2m18 ""; do
  echo "== variant $v"
  if [ -n "$v" ]; then export TF_LIB_PATH=$PWD/profiles/tools/_var/libtf_$v.so; else unset TF_LIB_PATH; fi
  python profiles/tools/gather_time.py 2>&1 | grep "sobel"
done
