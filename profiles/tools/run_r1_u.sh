#!/bin/bash
# GPU pass U (1 GPU): which part of the prologue reordering costs / pays: ordered (volatile) loads per group, round skipping
mkdir -p gpurun_out/u
O=gpurun_out/u
for lib in old p3n p3nb p3 p3b p3i p3x; do
  for wl in "water_24k 24" "water_192k 24"; do
    set -- $wl
    B200NB_LIBRARY=$PWD/scratch/lib_$lib.so timeout 300 python profiles/tools/kernel_sweep.py $1 $2 2>&1 | grep -v Warning | tee -a $O/sweep.txt
  done
done
