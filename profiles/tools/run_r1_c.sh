#!/bin/bash
# GPU pass C: second-generation force kernel (2 i-atoms x 1 j-atom per lane, LDGSTS staging): parity, bench, ncu
mkdir -p gpurun_out/c
O=gpurun_out/c
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --workload water_192k --no-cpu > $O/bench_192k.json 2> $O/bench_192k.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k python bench.py --workload water_192k --steps 3 --warmup 3 --no-cpu > $O/ncu_full_192k.log 2>&1
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench.json $O/bench_192k.json
