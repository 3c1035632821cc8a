#!/bin/bash
# GPU pass T of round 2 (1 GPU): register gather without predicates (rows end in dummy steps), 20 against 16 resident warps
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reflist.py tests/test_gpu_paths.py -q -x > $O/pytest_parity.txt 2>&1; tail -5 $O/pytest_parity.txt
for lib in default mb16; do
  for wl in water_24k water_192k water_1M; do
    L=""; [ $lib != default ] && L=scratch/lib_$lib.so
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 rf 2>&1 | grep -v Warning | tail -1 | sed "s/^/default rf /" | tee -a $O/sweep.txt
B200NB_LIBRARY=scratch/lib_mb16.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 rf 2>&1 | grep -v Warning | tail -1 | sed "s/^/mb16 rf /" | tee -a $O/sweep.txt
