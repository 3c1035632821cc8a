#!/bin/bash
# GPU pass P of round 2 (1 GPU): re-run of what pass O found broken (repartition API, one output buffer in the C++ benchmark
# driver) + the C++ benchmark driver at 96 k atoms
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 1500 python -m pytest tests/test_shim.py tests/test_gpu_domdec.py -q > $O/pytest_shim_dd.txt 2>&1; tail -6 $O/pytest_shim_dd.txt
OMP_PROC_BIND=spread OMP_PLACES=cores timeout 600 ./shim/_build/nbnxm_bench_gpu 32 pme 50 > $O/nbnxm_bench_gpu_96k_pme.json 2> $O/nbnxm_bench_gpu_96k.err; tail -1 $O/nbnxm_bench_gpu_96k_pme.json
OMP_PROC_BIND=spread OMP_PLACES=cores timeout 600 ./shim/_build/nbnxm_bench_gpu 32 rf 50 > $O/nbnxm_bench_gpu_96k_rf.json 2>> $O/nbnxm_bench_gpu_96k.err; tail -1 $O/nbnxm_bench_gpu_96k_rf.json
OMP_PROC_BIND=spread OMP_PLACES=cores timeout 600 ./shim/_build/nbnxm_bench_gpu 8 pme 200 > $O/nbnxm_bench_gpu_24k_pme.json 2>> $O/nbnxm_bench_gpu_96k.err; tail -1 $O/nbnxm_bench_gpu_24k_pme.json
tail -3 $O/nbnxm_bench_gpu_96k.err
