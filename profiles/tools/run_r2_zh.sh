#!/bin/bash
# GPU pass ZH of round 2 (1 GPU): unit tests of the repartitioning kernels, the same under compute-sanitizer memcheck, the DD suite
mkdir -p gpurun_out/r2zh
O=gpurun_out/r2zh
timeout 900 python -m pytest tests/test_gpu_partition.py -q > $O/pytest_partition.txt 2>&1; tail -5 $O/pytest_partition.txt
timeout 1200 compute-sanitizer --tool memcheck --log-file $O/memcheck_partition.log python -m pytest tests/test_gpu_partition.py -q -x > $O/memcheck_partition.pytest.txt 2>&1
echo "memcheck partition kernels: $(tail -1 $O/memcheck_partition.pytest.txt) | $(grep 'ERROR SUMMARY' $O/memcheck_partition.log | tail -1)"
timeout 1500 python -m pytest tests/test_gpu_domdec.py tests/test_gpu_domdec_ipc.py -q > $O/pytest_dd.txt 2>&1; tail -3 $O/pytest_dd.txt
