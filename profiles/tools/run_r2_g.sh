#!/bin/bash
# GPU pass G of round 2 (1 GPU): the whole -m gpu suite with the tightened energy bars, the multi-process CUDA-IPC test and the
# tests of the previously untested paths
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -40 $O/pytest_gpu.txt
