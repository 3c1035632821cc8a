#!/bin/bash
# GPU pass E of round 2 (1 GPU): the compiled reference-side binding -- the reference's nblib force tests with useGpu = true
# through shim/_build -- and the reference-built-list parity tests
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 300 ./shim/_build/nblib_gpu_test > $O/nblib_gpu_test.json 2> $O/nblib_gpu_test.err; echo "nblib_gpu_test rc=$?"; head -c 400 $O/nblib_gpu_test.json; tail -5 $O/nblib_gpu_test.err
timeout 900 python -m pytest tests/test_shim.py tests/test_gpu_reflist.py -q > $O/pytest_shim_reflist.txt 2>&1; tail -25 $O/pytest_shim_reflist.txt
