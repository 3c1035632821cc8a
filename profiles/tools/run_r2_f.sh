#!/bin/bash
# GPU pass F of round 2 (1 GPU): shim tests (reference nblib force tests with useGpu = true), the new bench.py line (water_1M
# default + water_24k secondary + search / sustained blocks), reference arm, launch list of a re-search at 24 k
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 300 ./shim/_build/nblib_gpu_test > $O/nblib_gpu_test.json 2> $O/nblib_gpu_test.err; echo "nblib_gpu_test rc=$?"; tail -3 $O/nblib_gpu_test.err
timeout 900 python -m pytest tests/test_shim.py tests/test_gpu_reflist.py -q > $O/pytest_shim_reflist.txt 2>&1; tail -8 $O/pytest_shim_reflist.txt
timeout 900 python bench.py --steps 50 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err; head -c 3000 $O/bench_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?"; head -c 1500 $O/bench_ref_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_24k.csv python bench.py --workload water_24k --secondary "" --steps 3 --warmup 3 --no-cpu --no-sustained > $O/ncu_launches.log 2>&1
ls -la $O
