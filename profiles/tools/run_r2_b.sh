#!/bin/bash
# GPU pass B of round 2 (1 GPU): why the re-laid-out kernel does not convert its 27 % fewer instructions into time:
# arithmetic ceiling of the step without memory (pairmath_bench), register / occupancy variants with the i-atoms in shared
# memory, cold vs warm L2, ncu of the 24-warp variant
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
./scratch/pairmath_bench > $O/pairmath_bench.txt 2>&1; cat $O/pairmath_bench.txt
for lib in s16 s20 s24 s32 m20; do
  for wl in water_24k water_192k water_1M; do
    B200NB_LIBRARY=$PWD/scratch/lib_$lib.so timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
  NOFLUSH=1 B200NB_LIBRARY=$PWD/scratch/lib_$lib.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 ewald 2>&1 | grep -v Warning | sed "s/^/$lib /" | tee -a $O/sweep.txt
done
B200NB_LIBRARY=$PWD/scratch/lib_s24.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k_s24 python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu > $O/ncu_full_192k.log 2>&1
ncu -i $O/prof_force_192k_s24.ncu-rep --page raw --csv > $O/prof_force_192k_s24_raw.csv 2>/dev/null; python profiles/tools/ncu_summary.py $O/prof_force_192k_s24_raw.csv > $O/ncu_k_force_water192k_s24.txt 2>&1; cat $O/ncu_k_force_water192k_s24.txt
# same capture with warm caches (what the kernel sees inside a step: xq just written, lj prefetched)
B200NB_LIBRARY=$PWD/scratch/lib_s24.so timeout 900 ncu --set full --cache-control none --clock-control none -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k_s24_warm python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu --no-flush > $O/ncu_full_192k_warm.log 2>&1
ncu -i $O/prof_force_192k_s24_warm.ncu-rep --page raw --csv > $O/prof_force_192k_s24_warm_raw.csv 2>/dev/null; python profiles/tools/ncu_summary.py $O/prof_force_192k_s24_warm_raw.csv > $O/ncu_k_force_water192k_s24_warm.txt 2>&1; cat $O/ncu_k_force_water192k_s24_warm.txt
ls -la $O
