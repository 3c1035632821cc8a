#!/bin/bash
# GPU pass ZQ of round 2 (1 GPU): compute-sanitizer memcheck + racecheck on the final tree: single-domain parity at 3 k atoms (half-entry
# packing, stable ordering, register-gather kernel), dynamic pruning (rolling parts), the triclinic cell
mkdir -p gpurun_out/r2zq
O=gpurun_out/r2zq
T="tests/test_gpu_parity.py"
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --log-file $O/${tool}_final.log python -m pytest $T -q -x -k "water_3k or triclinic or dynamic" > $O/${tool}_final.pytest.txt 2>&1
  echo "$tool: $(tail -1 $O/${tool}_final.pytest.txt) | $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/${tool}_final.log | tail -1)"
done
