#!/bin/bash
# GPU pass Y (1 GPU): persistent warps walking the size-sorted entries in boustrophedon order (B200NB_FORCE_ROUNDS = longest
# list, in waves of resident warps, that runs this way; 0 = one entry per warp as before)
mkdir -p gpurun_out/y
O=gpurun_out/y
for rounds in 0 3 8 100; do
  for wl in "water_24k 16,24,32,0" "water_96k 24,32,0" "water_192k 24,0"; do
    set -- $wl
    B200NB_FORCE_ROUNDS=$rounds B200NB_LIBRARY=$PWD/scratch/lib_persist.so timeout 300 python profiles/tools/kernel_sweep.py $1 $2 2>&1 | grep -v Warning | sed "s/^/rounds=$rounds /" | tee -a $O/sweep.txt
  done
done
B200NB_LIBRARY=$PWD/scratch/lib_persist.so timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_gpu.txt
