#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>: builds scratch/lib_<name>.so from the current sources (tuning experiments)
set -e
name=$1; shift
mkdir -p scratch/obj_$name
for f in b200nb force; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Iinclude -Igmxapi_b200/csrc "$@" -c gmxapi_b200/csrc/$f.cu -o scratch/obj_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/lib_$name.so scratch/obj_$name/*.o
echo built scratch/lib_$name.so
