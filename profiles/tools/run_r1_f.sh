#!/bin/bash
# GPU pass F: pitch-addressed packed list (2-level prologue), L2 prefetch in k_step_begin, replicated output accumulators
mkdir -p gpurun_out/f
O=gpurun_out/f
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
( timeout 300 python profiles/tools/kernel_sweep.py water_24k 12,16,24,32
  for v in i1b32 i2b28 i1b24 i2b24; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_24k 12,16,24; done
  B200NB_LIBRARY=$PWD/scratch/lib_w2i2.so timeout 300 python profiles/tools/kernel_sweep.py water_24k 16,24
  timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24,32
  for v in i1b32 i2b28 i2b24; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24; done ) > $O/sweep.txt 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --workload water_192k --no-cpu > $O/bench_192k.json 2> $O/bench_192k.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/bench.json $O/bench_192k.json; cat $O/sweep.txt
