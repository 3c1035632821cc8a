#!/bin/bash
# GPU pass N of round 2 (1 GPU): `ncu --set full` of the force kernel at 24 k, 192 k and 1 M atoms (roofline.traffic, pipe
# utilisation, source-level stall table), then compute-sanitizer (pass L)
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
for wl in water_24k water_192k water_1M; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_$wl python bench.py --workload $wl --secondary "" --steps 3 --warmup 3 --no-cpu --no-search --no-sustained > $O/ncu_$wl.log 2>&1
  ncu -i $O/prof_$wl.ncu-rep --page raw --csv > $O/raw_$wl.csv 2>/dev/null
  python profiles/tools/ncu_summary.py $O/raw_$wl.csv > $O/ncu_k_force_$wl.txt 2>&1
  ncu -i $O/prof_$wl.ncu-rep --page source --csv > $O/src_$wl.csv 2>/dev/null
  rm -f $O/prof_$wl.ncu-rep
  grep "gpu__time_duration\|fma_cycles_active\|issue_active\|dram__bytes_read" $O/ncu_k_force_$wl.txt
done
bash profiles/tools/run_r2_l.sh
