#!/bin/bash
# GPU pass D: j-atom-packed list, fused step kernels, zero-copy host path: parity, bench, variant sweep, ncu
mkdir -p gpurun_out/d
O=gpurun_out/d
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --workload water_192k --no-cpu > $O/bench_192k.json 2> $O/bench_192k.err
( timeout 300 python profiles/tools/kernel_sweep.py water_24k 8,12,16,24,32
  for v in w1 w4 w2ilp1 w1ilp1; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_24k 8,16,32; done
  timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,32
  B200NB_LIBRARY=$PWD/scratch/lib_w1.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,32
  B200NB_LIBRARY=$PWD/scratch/lib_w4.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,32 ) > $O/sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench.json $O/bench_192k.json; cat $O/sweep.txt
