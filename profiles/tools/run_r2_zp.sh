#!/bin/bash
# GPU pass ZP of round 2 (1 GPU): triclinic cells (b200nb_set_box_triclinic) against the triclinic mode of the oracle and the reference's
# own outputs; the single-domain parity file around it
mkdir -p gpurun_out/r2zp
O=gpurun_out/r2zp
timeout 1200 python -m pytest tests/test_gpu_parity.py -q > $O/pytest_parity.txt 2>&1; tail -15 $O/pytest_parity.txt
