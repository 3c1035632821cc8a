#!/bin/bash
# GPU pass ZZF of round 2 (1 GPU): rvdw < rcoulomb in the free-energy kernel (new test) + the LJ-PME tests on the same kernel
mkdir -p gpurun_out/r2zzf
timeout 15 python -m pytest tests/test_gpu_fep.py -q -k "twin_range or ljpme" > gpurun_out/r2zzf/pytest_fep_twin.txt 2>&1; tail -30 gpurun_out/r2zzf/pytest_fep_twin.txt
