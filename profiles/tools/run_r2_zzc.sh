#!/bin/bash
# GPU pass ZZC of round 2 (1 GPU): LJ potential switch in the free-energy kernel
mkdir -p gpurun_out/r2zzc
timeout 300 python -m pytest tests/test_gpu_fep.py -q -x -k "potential_switch or refuses" > gpurun_out/r2zzc/pytest_fep_pswitch.txt 2>&1; tail -25 gpurun_out/r2zzc/pytest_fep_pswitch.txt
