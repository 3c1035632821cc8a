#!/bin/bash
# GPU pass ZF of round 2 (1 GPU): ncu --set full of k_search and k_pack at 1 M atoms (steady-state search step)
mkdir -p gpurun_out/r2zf
O=gpurun_out/r2zf
for k in k_search k_pack; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 2 -c 1 -o $O/prof_$k python bench.py --steps 2 --warmup 1 --no-cpu --no-sustained --secondary "" > $O/ncu_$k.log 2>&1
  ncu -i $O/prof_$k.ncu-rep --page raw --csv > $O/raw_$k.csv 2>/dev/null
  python profiles/tools/ncu_summary.py $O/raw_$k.csv > $O/ncu_${k}_water_1M.txt 2>&1; cat $O/ncu_${k}_water_1M.txt | cut -c1-150
  ncu -i $O/prof_$k.ncu-rep --page source --csv > $O/src_$k.csv 2>/dev/null
  rm -f $O/prof_$k.ncu-rep
done
