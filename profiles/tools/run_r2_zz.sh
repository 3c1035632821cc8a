#!/bin/bash
# GPU pass ZZ of round 2 (1 GPU): the device-built perturbed pair list incl. triclinic cells
mkdir -p gpurun_out/r2zz
timeout 600 python -m pytest tests/test_gpu_fep.py -q -x > gpurun_out/r2zz/pytest_fep.txt 2>&1; tail -30 gpurun_out/r2zz/pytest_fep.txt
