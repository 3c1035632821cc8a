#!/bin/bash
# GPU pass ZD of round 2 (1 GPU): N-D repartitioning on the device against the numpy restatement; the whole DD suite, twice (the
# in-process ranks push from the main stream; the two-process IPC test runs the separate push branch)
mkdir -p gpurun_out/r2zd
O=gpurun_out/r2zd
for i in 1 2; do
timeout 1500 python -m pytest tests/test_gpu_domdec.py tests/test_gpu_domdec_ipc.py -q > $O/pytest_dd_$i.txt 2>&1; tail -4 $O/pytest_dd_$i.txt
done
