#!/bin/bash
# GPU pass X (1 GPU): ring-staged force kernel (constant shared memory, entries up to 64 tiles) against whole-entry staging
mkdir -p gpurun_out/x
O=gpurun_out/x
for lib in noring ring ring12; do
  for wl in "water_24k 16,24,32" "water_96k 24,32,48,64" "water_192k 24,32,48,64"; do
    set -- $wl
    mts=$2; [ $lib = noring ] && mts=$(echo $2 | sed 's/,48,64//')
    B200NB_LIBRARY=$PWD/scratch/lib_$lib.so timeout 300 python profiles/tools/kernel_sweep.py $1 $mts 2>&1 | grep -v Warning | tee -a $O/sweep.txt
  done
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_gpu.txt
