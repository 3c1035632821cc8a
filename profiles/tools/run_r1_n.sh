#!/bin/bash
# GPU pass N (1 GPU): fused DD kernels + CUDA-graph step replay: tests, bench, ncu of the current default force kernel
mkdir -p gpurun_out/n
O=gpurun_out/n
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log; grep -E "^E  " $O/pytest_gpu.log | head
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --workload water_192k --no-cpu > $O/bench_192k.json 2> $O/bench_192k.err; cat $O/bench_192k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1
