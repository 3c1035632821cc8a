#!/bin/bash
# GPU pass ZS of round 2 (1 GPU): triclinic cells incl. the shear at the limits of check_box, where pairs sit two box vectors away along x
mkdir -p gpurun_out/r2zs
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k triclinic > gpurun_out/r2zs/pytest_triclinic.txt 2>&1; tail -12 gpurun_out/r2zs/pytest_triclinic.txt
