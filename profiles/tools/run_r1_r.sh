#!/bin/bash
# GPU pass R (1 GPU): entry-length sweep with size-sorted entries; launch list and full ncu capture of the final kernel
mkdir -p gpurun_out/r
O=gpurun_out/r
( timeout 300 python profiles/tools/kernel_sweep.py water_24k 10,12,14,16,20,24,32
  timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24,32
  timeout 300 python profiles/tools/kernel_sweep.py water_96k 16,24 ) > $O/sweep.txt 2>&1
cat $O/sweep.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu > $O/ncu_full_192k.log 2>&1
