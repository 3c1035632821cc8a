#!/bin/bash
# GPU pass ZZB of round 2 (1 GPU): the files touched since the last full -m gpu pass (ZW): bonded, free-energy, shim tests; smoke
mkdir -p gpurun_out/r2zzb
O=gpurun_out/r2zzb
timeout 900 python -m pytest tests/test_gpu_bonded.py tests/test_gpu_fep.py tests/test_shim.py tests/test_gpu_paths.py -m gpu -q > $O/pytest_touched.txt 2>&1; tail -6 $O/pytest_touched.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt | cut -c1-300
