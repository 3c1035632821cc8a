#!/bin/bash
# GPU pass ZN of round 2 (1 GPU): execution order with long and short entries interleaved (B200NB_ORDER_INTERLEAVE=1) against strictly
# descending sizes: does a wave of equally long entries run as a convoy?
mkdir -p gpurun_out/r2zn
O=gpurun_out/r2zn
for il in 0 1; do
  for cfg in "water_24k ewald" "water_96k ewald" "water_192k ewald" "water_1M ewald"; do
    set -- $cfg
    B200NB_ORDER_INTERLEAVE=$il timeout 300 python profiles/tools/kernel_sweep.py $1 0 $2 2>&1 | grep -v Warning | tail -1 | sed "s/^/interleave=$il /" | tee -a $O/sweep.txt
  done
done
B200NB_ORDER_INTERLEAVE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "pairs_forces or dynamic" > $O/pytest.txt 2>&1; tail -2 $O/pytest.txt
