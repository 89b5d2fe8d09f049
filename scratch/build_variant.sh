#!/bin/bash
# usage: scratch/build_variant.sh NAME -DFOO=1 ...   -> scratch/_var/libtf_NAME.so
set -e
name=$1; shift
mkdir -p scratch/_var/obj_$name
for f in tobac_flow_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o scratch/_var/obj_$name/$b.o &
done
wait
nvcc -shared -o scratch/_var/libtf_$name.so scratch/_var/obj_$name/*.o
echo scratch/_var/libtf_$name.so
