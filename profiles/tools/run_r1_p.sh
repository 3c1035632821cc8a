#!/bin/bash
# GPU pass P (1 GPU): FMA-pipe trims (beta folded into the denominator, FFMA2 i-accumulation) + CTA-size sweep
mkdir -p gpurun_out/p
O=gpurun_out/p
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > $O/pytest_parity.log 2>&1; tail -2 $O/pytest_parity.log
( timeout 300 python profiles/tools/kernel_sweep.py water_24k 12,16,20,24
  for v in w2b16 w4b8 w8b4; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_24k 12,16,24; done
  timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24,32
  for v in w2b16 w4b8; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24; done
  timeout 300 python profiles/tools/kernel_sweep.py water_24k 16 rf ) > $O/sweep.txt 2>&1
cat $O/sweep.txt
