#!/bin/bash
# build_force_variant.sh <name> <force.cu path> <extra nvcc flags...>: scratch/lib_<name>.so = current b200nb.o + a variant of force.cu
set -e
name=$1; src=$2; shift 2
mkdir -p scratch
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Iinclude -Igmxapi_b200/csrc"
[ -f gmxapi_b200/csrc/b200nb.o ] || make -C gmxapi_b200/csrc
$NV "$@" -c $src -o scratch/force_$name.o
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/lib_$name.so gmxapi_b200/csrc/b200nb.o gmxapi_b200/csrc/dd_partition.o scratch/force_$name.o
echo built scratch/lib_$name.so
