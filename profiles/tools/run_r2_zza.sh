#!/bin/bash
# GPU pass ZZA of round 2 (1 GPU): gmx::GpuBonded on the C ABI, driven as do_force() drives it (shim/gpubonded_test.cpp)
mkdir -p gpurun_out/r2zza
timeout 120 shim/_build/gpubonded_test > gpurun_out/r2zza/gpubonded_test.json 2> gpurun_out/r2zza/gpubonded_test.err; echo "rc=$?"; cat gpurun_out/r2zza/gpubonded_test.json; tail -5 gpurun_out/r2zza/gpubonded_test.err
