#!/bin/bash
# GPU pass V (1 GPU): where the pair loop's time goes -- diagnostic builds (wrong results) without the j-force scatter, without
# the whole j-force reduction, without MUFU; and the reaction-field kernel (13 fewer packed FMAs per tile) as the FMA slope
mkdir -p gpurun_out/v
O=gpurun_out/v
for lib in old d_nored d_nojf d_nomufu d_nomufu_nojf; do
  for wl in "water_24k 24 ewald" "water_192k 24 ewald" "water_192k 24 rf"; do
    set -- $wl
    B200NB_LIBRARY=$PWD/scratch/lib_$lib.so timeout 300 python profiles/tools/kernel_sweep.py $1 $2 $3 2>&1 | grep -v Warning | sed "s/^/$3 /" | tee -a $O/sweep.txt
  done
done
