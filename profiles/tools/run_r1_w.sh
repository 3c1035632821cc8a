#!/bin/bash
# GPU pass W (1 GPU): full GPU test suite with the LJ-modifier kernels; single-GPU bench lines of the larger BASELINE configurations
mkdir -p gpurun_out/w
O=gpurun_out/w
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_gpu.txt
for wl in "water_96k ewald" "water_1M ewald" "water_1.5M rf"; do
  set -- $wl
  timeout 600 python bench.py --no-cpu --workload $1 --eel $2 --steps 50 --warmup 5 > $O/bench_$1_$2.json 2> $O/bench_$1_$2.err
  tail -c 1500 $O/bench_$1_$2.json; tail -3 $O/bench_$1_$2.err
done
