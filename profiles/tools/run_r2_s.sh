#!/bin/bash
# GPU pass S of round 2 (1 GPU): j-atoms gathered into registers one step ahead (loads pinned above the deferred j-force red)
# against the cp.async ring: the L1TEX data pipe was at 86 % with the ring (r2r ncu)
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reflist.py tests/test_gpu_paths.py -q -x > $O/pytest_parity.txt 2>&1; tail -5 $O/pytest_parity.txt
for lib in default u2 ring; do
  for wl in water_24k water_192k water_1M; do
    L=""; [ $lib != default ] && L=scratch/lib_$lib.so
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 rf 2>&1 | grep -v Warning | tail -1 | sed "s/^/default rf /" | tee -a $O/sweep.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_1M python bench.py --workload water_1M --secondary "" --steps 3 --warmup 3 --no-cpu --no-search --no-sustained > $O/ncu_full_1M.log 2>&1
ncu -i $O/prof_force_1M.ncu-rep --page raw --csv > $O/prof_force_1M_raw.csv 2>/dev/null; timeout 300 python profiles/tools/ncu_summary.py $O/prof_force_1M_raw.csv > $O/ncu_k_force_water1M.txt 2>&1; cat $O/ncu_k_force_water1M.txt
grep -o "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed[^,]*" $O/prof_force_1M_raw.csv | head -1
ncu -i $O/prof_force_1M.ncu-rep --page source --csv > $O/src_sass.csv 2>/dev/null
rm -f $O/prof_force_1M.ncu-rep
