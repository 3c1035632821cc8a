#!/bin/bash
# GPU pass ZL of round 2 (1 GPU): the j-force red as v2 + scalar (3 float adds in L2, 2 requests) against v4 (4 adds, 1 request) on the
# register-gather kernel: the L2 reduction rate (~0.6 T adds/s) equals the kernel time of the reaction-field kernel with v4
mkdir -p gpurun_out/r2zl
O=gpurun_out/r2zl
for lib in default split; do
  for cfg in "water_1M ewald" "water_1M rf" "water_192k ewald" "water_24k ewald"; do
    set -- $cfg
    L=""; [ $lib != default ] && L=scratch/lib_$lib.so
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $1 0 $2 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib $2 /" | tee -a $O/sweep.txt
  done
done
