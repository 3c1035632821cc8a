#!/bin/bash
# GPU pass T (1 GPU): two-level prologue (ordered loads), rotated pair loop, warps per CTA: kernel-only A/B + parity
mkdir -p gpurun_out/t
O=gpurun_out/t
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/gpu.txt
for lib in old p2 p2rot p2rot_w2 p2rot_w4 p2_w4; do
  for wl in "water_24k 16,24" "water_192k 24"; do
    set -- $wl
    B200NB_LIBRARY=$PWD/scratch/lib_$lib.so timeout 300 python profiles/tools/kernel_sweep.py $1 $2 2>&1 | grep -v Warning | tee -a $O/sweep.txt
  done
done
for lib in p2rot p2; do
  echo "== parity $lib" | tee -a $O/parity.txt
  B200NB_LIBRARY=$PWD/scratch/lib_$lib.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee -a $O/parity.txt
done
