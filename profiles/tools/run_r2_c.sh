#!/bin/bash
# GPU pass C of round 2 (1 GPU): the force kernel with the per-lane cp.async j-atom ring (gather NB_RING-1 steps ahead, nothing
# in flight held in registers): parity, sweep of ring depth / resident warps on 24 k, 192 k, 1 M atoms, full ncu at 192 k with
# the source page
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > $O/pytest_parity.txt 2>&1; tail -5 $O/pytest_parity.txt
for lib in default m20 m24 r3 r6 m20r6; do
  for wl in water_24k water_192k water_1M; do
    if [ $lib = default ]; then L=$PWD/gmxapi_b200/libb200nb.so; else L=$PWD/scratch/lib_$lib.so; fi
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
B200NB_LIBRARY=$PWD/gmxapi_b200/libb200nb.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 rf 2>&1 | grep -v Warning | tail -1 | sed "s/^/default rf /" | tee -a $O/sweep.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu > $O/ncu_full_192k.log 2>&1
ncu -i $O/prof_force_192k.ncu-rep --page raw --csv > $O/prof_force_192k_raw.csv 2>/dev/null; timeout 300 python profiles/tools/ncu_summary.py $O/prof_force_192k_raw.csv > $O/ncu_k_force_water192k.txt 2>&1; cat $O/ncu_k_force_water192k.txt
ncu -i $O/prof_force_192k.ncu-rep --page source --csv > $O/src_sass.csv 2>/dev/null
rm -f $O/prof_force_192k.ncu-rep
ls -la $O
