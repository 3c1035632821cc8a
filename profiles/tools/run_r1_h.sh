#!/bin/bash
# GPU pass H: persistent pipelined force kernel
mkdir -p gpurun_out/h
O=gpurun_out/h
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( timeout 300 python profiles/tools/kernel_sweep.py water_24k 8,12,16,20,24
  for v in pi1 pw2 pw8; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_24k 12,16,24; done
  timeout 300 python profiles/tools/kernel_sweep.py water_192k 12,16,24
  for v in pi1 pw2; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24; done ) > $O/sweep.txt 2>&1
cat $O/sweep.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1
