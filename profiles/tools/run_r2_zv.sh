#!/bin/bash
# GPU pass ZV of round 2 (1 GPU): the listed-interaction (bonded) kernel against the restated and the committed reference outputs
mkdir -p gpurun_out/r2zv
timeout 600 python -m pytest tests/test_gpu_bonded.py -q -x > gpurun_out/r2zv/pytest_bonded.txt 2>&1; tail -40 gpurun_out/r2zv/pytest_bonded.txt
