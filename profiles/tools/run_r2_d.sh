#!/bin/bash
# GPU pass D of round 2 (1 GPU): whole -m gpu suite (new: reference-built list through b200nb_upload_pairlist), force kernel with
# the index stream riding the cp.async ring and 20 resident warps, ncu + source page at 192 k
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -15 $O/pytest_gpu.txt
for wl in water_24k water_192k water_1M; do
  timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | tee -a $O/sweep.txt
done
timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 rf 2>&1 | grep -v Warning | tail -1 | sed "s/^/rf /" | tee -a $O/sweep.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu > $O/ncu_full_192k.log 2>&1
ncu -i $O/prof_force_192k.ncu-rep --page raw --csv > $O/prof_force_192k_raw.csv 2>/dev/null; timeout 300 python profiles/tools/ncu_summary.py $O/prof_force_192k_raw.csv > $O/ncu_k_force_water192k.txt 2>&1; cat $O/ncu_k_force_water192k.txt
ncu -i $O/prof_force_192k.ncu-rep --page source --csv > $O/src_sass.csv 2>/dev/null
rm -f $O/prof_force_192k.ncu-rep
ls -la $O
