#!/bin/bash
# GPU pass ZT of round 2 (1 GPU): the perturbed-pair (free-energy) kernel against the restated and the committed reference kernel outputs
mkdir -p gpurun_out/r2zt
timeout 900 python -m pytest tests/test_gpu_fep.py -q > gpurun_out/r2zt/pytest_fep.txt 2>&1; tail -25 gpurun_out/r2zt/pytest_fep.txt
