#!/bin/bash
# GPU pass L (1 GPU): full GPU test suite with the two-stream DD step
mkdir -p gpurun_out/l
O=gpurun_out/l
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log; grep -E "^E  " $O/pytest_gpu.log | head
