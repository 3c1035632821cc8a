#!/bin/bash
# GPU pass ZW of round 2 (1 GPU): the whole -m gpu suite, smoke and a short bench on the final tree (with the free-energy list /
# kernel and the bonded kernel), then compute-sanitizer memcheck + racecheck on the new kernels' tests
mkdir -p gpurun_out/r2zw
O=gpurun_out/r2zw
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -5 $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'search', d['search']['search_ms'])"
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 400 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bonded.py "tests/test_gpu_fep.py::test_fep_list_built_on_the_device[water_3k-20-1.0]" "tests/test_gpu_fep.py::test_fep_kernel_matches_oracle_and_reference[sc1coul-ewald]" -q -x > $O/memcheck_fep_bonded.txt 2>&1; echo "memcheck rc=$?"; tail -4 $O/memcheck_fep_bonded.txt
timeout 400 $SAN --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_bonded.py::test_bonded_kernel_matches_oracle_and_reference[rect]" "tests/test_gpu_fep.py::test_fep_list_built_on_the_device[water_3k-20-1.0]" -q -x > $O/racecheck_fep_bonded.txt 2>&1; echo "racecheck rc=$?"; tail -4 $O/racecheck_fep_bonded.txt
