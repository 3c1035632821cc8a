#!/bin/bash
# GPU pass ZY of round 2 (1 GPU): bonded and free-energy kernels as part of the captured step (b200nb_bonded_in_step, b200nb_fep_in_step)
mkdir -p gpurun_out/r2zy
timeout 600 python -m pytest tests/test_gpu_bonded.py tests/test_gpu_fep.py -q -x -k "bonded_kernel_matches or list_built" > gpurun_out/r2zy/pytest_in_step.txt 2>&1; tail -30 gpurun_out/r2zy/pytest_in_step.txt
