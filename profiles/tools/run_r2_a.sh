#!/bin/bash
# GPU pass A of round 2 (1 GPU): parity of the re-laid-out force kernel (one j-atom per lane, 16 j x 2 i-halves), sweep of its
# build variants (registers / warps per CTA) on 24 k, 192 k and 1 M atoms, one full ncu capture at 192 k
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/pytest_parity.txt 2>&1; tail -5 $O/pytest_parity.txt
for lib in default m20 m24 w2 w4; do
  for wl in "water_24k 0" "water_192k 0" "water_1M 0"; do
    set -- $wl
    if [ $lib = default ]; then L=$PWD/gmxapi_b200/libb200nb.so; else L=$PWD/scratch/lib_$lib.so; fi
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $1 $2 ewald 2>&1 | grep -v Warning | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
B200NB_LIBRARY=$PWD/gmxapi_b200/libb200nb.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 rf 2>&1 | grep -v Warning | sed "s/^/default rf /" | tee -a $O/sweep.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu > $O/ncu_full_192k.log 2>&1
ncu -i $O/prof_force_192k.ncu-rep --page raw --csv > $O/prof_force_192k_raw.csv 2>/dev/null; timeout 300 python profiles/tools/ncu_summary.py $O/prof_force_192k_raw.csv > $O/ncu_k_force_water192k.txt 2>&1; cat $O/ncu_k_force_water192k.txt
timeout 300 python profiles/tools/sustained.py water_192k 3.0 > $O/sustained_192k.txt 2>&1; cat $O/sustained_192k.txt
ls -la $O
