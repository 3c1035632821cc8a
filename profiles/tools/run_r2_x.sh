#!/bin/bash
# GPU pass V of round 2 (1 GPU): what the divergent memory instructions of a step cost (diagnostic builds: no LJ gather, no red)
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
for lib in default a2 a2mb18; do
  for wl in water_24k water_192k water_1M; do
    L=""; [ $lib != default ] && L=scratch/lib_$lib.so
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
