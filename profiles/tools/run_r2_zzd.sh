#!/bin/bash
# GPU pass ZZD of round 2 (1 GPU): free-energy and bonded tests after the list arrays became persistent (no cudaMalloc in a rebuild), shim bonded test
mkdir -p gpurun_out/r2zzd
timeout 600 python -m pytest tests/test_gpu_fep.py tests/test_gpu_bonded.py "tests/test_shim.py::test_reference_gpubonded_interface_on_b200nb" -q -x > gpurun_out/r2zzd/pytest_fep_bonded.txt 2>&1; tail -8 gpurun_out/r2zzd/pytest_fep_bonded.txt
