#!/bin/bash
# GPU pass L of round 2 (1 GPU): compute-sanitizer (memcheck, racecheck, synccheck) on the single-domain parity test at 3 k atoms
# and on the 2-rank decomposed step (peer-window flag protocol, last-CTA counters, graph replay) at 24 k atoms
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
T1='tests/test_gpu_parity.py::test_pairs_forces_energies[water_3k-CoulombType.Pme]'
T2='tests/test_gpu_domdec.py::test_domain_decomposition_matches_single_domain[water_24k-2-CoulombType.Pme-True]'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file $O/${tool}_parity_water3k.log python -m pytest "$T1" -q -x > $O/${tool}_parity_water3k.pytest.txt 2>&1
  echo "$tool parity: $(tail -1 $O/${tool}_parity_water3k.pytest.txt) | $(grep -c 'ERROR SUMMARY' $O/${tool}_parity_water3k.log) summaries: $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/${tool}_parity_water3k.log | tail -1)"
done
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file $O/${tool}_dd2_water24k.log python -m pytest "$T2" -q -x > $O/${tool}_dd2_water24k.pytest.txt 2>&1
  echo "$tool dd2: $(tail -1 $O/${tool}_dd2_water24k.pytest.txt) | $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/${tool}_dd2_water24k.log | tail -1)"
done
ls -la $O
