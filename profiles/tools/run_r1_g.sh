#!/bin/bash
# GPU pass G: in-step kernel timing sweep + full tests
mkdir -p gpurun_out/g
O=gpurun_out/g
( timeout 300 python profiles/tools/kernel_sweep.py water_24k 12,16,24,32
  for v in i1b32 i2b28 i2b24; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_24k 12,16,24; done
  timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24,32
  for v in i1b32 i2b28; do B200NB_LIBRARY=$PWD/scratch/lib_$v.so timeout 300 python profiles/tools/kernel_sweep.py water_192k 16,24; done ) > $O/sweep.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
tail -3 $O/pytest_gpu.log; cat $O/bench.json; cat $O/sweep.txt
