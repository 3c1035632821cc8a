#!/bin/bash
# GPU pass U of round 2 (1 GPU): register gather + L1 / L2 prefetch of the j-atoms two steps ahead
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
for lib in default pf1 pf2; do
  for wl in water_24k water_192k water_1M; do
    L=""; [ $lib != default ] && L=scratch/lib_$lib.so
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
