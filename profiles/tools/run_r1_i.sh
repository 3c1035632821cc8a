#!/bin/bash
# GPU pass I: one-entry-per-warp kernel at 32 warps/SM (defaults), peer-window DD step on one GPU (loopback ranks)
mkdir -p gpurun_out/i
O=gpurun_out/i
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -25 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json
timeout 600 python bench.py --workload water_192k --no-cpu > $O/bench_192k.json 2> $O/bench_192k.err; cat $O/bench_192k.json
