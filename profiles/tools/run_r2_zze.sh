#!/bin/bash
# GPU pass ZZE of round 2 (1 GPU): LJ-PME in the free-energy kernel (2 new tests), then the whole free-energy file on the changed kernel
mkdir -p gpurun_out/r2zze
timeout 25 python -m pytest tests/test_gpu_fep.py -q -k ljpme > gpurun_out/r2zze/pytest_fep_ljpme.txt 2>&1; tail -30 gpurun_out/r2zze/pytest_fep_ljpme.txt
timeout 25 python -m pytest tests/test_gpu_fep.py -q -x > gpurun_out/r2zze/pytest_fep_all.txt 2>&1; tail -5 gpurun_out/r2zze/pytest_fep_all.txt
