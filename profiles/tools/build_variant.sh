#!/bin/bash
# usage: profiles/tools/build_variant.sh NAME -DFOO=1 ...   -> profiles/tools/_var/libtf_NAME.so
set -e
name=$1; shift
mkdir -p profiles/tools/_var/obj_$name
for f in tobac_flow_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o profiles/tools/_var/obj_$name/$b.o &
done
wait
nvcc -shared -o profiles/tools/_var/libtf_$name.so profiles/tools/_var/obj_$name/*.o
echo profiles/tools/_var/libtf_$name.so
