#!/bin/bash
# GPU pass Z (1 GPU): bench lines (24k with the CPU reference beside it, 96k, 192k, 1M, 1.5M RF), launch list and full ncu capture of the ring kernel
mkdir -p gpurun_out/z
O=gpurun_out/z
timeout 900 python bench.py > $O/bench_water24k.json 2> $O/bench_water24k.err; tail -c 600 $O/bench_water24k.json
timeout 600 python bench.py --impl reference --steps 200 --warmup 3 > $O/bench_reference_water24k.json 2> $O/bench_reference.err; tail -c 400 $O/bench_reference_water24k.json
for wl in "water_96k ewald" "water_192k ewald" "water_1M ewald" "water_1.5M rf"; do
  set -- $wl
  timeout 600 python bench.py --no-cpu --workload $1 --eel $2 --steps 50 --warmup 5 > $O/bench_$1_$2.json 2> $O/bench_$1_$2.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_water24k.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_24k python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full_24k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu > $O/ncu_full_192k.log 2>&1
ls -la $O
