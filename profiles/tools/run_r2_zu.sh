#!/bin/bash
# GPU pass ZU of round 2 (1 GPU): the perturbed pair list built on the device (b200nb_fep_build_list) against the oracle's list
mkdir -p gpurun_out/r2zu
timeout 600 python -m pytest tests/test_gpu_fep.py -q -x -k "list_built" > gpurun_out/r2zu/pytest_fep_list.txt 2>&1; tail -30 gpurun_out/r2zu/pytest_fep_list.txt
