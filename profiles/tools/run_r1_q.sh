#!/bin/bash
# GPU pass Q (1 GPU): size-sorted packed entries + programmatic dependent launch of the force kernel
mkdir -p gpurun_out/q
O=gpurun_out/q
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log; grep -E "^E  " $O/pytest_gpu.log | head
for pdl in 1 0; do
B200NB_PDL=$pdl timeout 600 python bench.py --no-cpu > $O/bench_pdl$pdl.json 2> $O/bench_pdl$pdl.err
python - <<PY
import json
d=json.loads(open("$O/bench_pdl$pdl.json").read().strip().splitlines()[-1]); print("PDL=$pdl step %.2f us  e2e %.2f us  kernel in step %.2f us (alone cold %.2f)"%(d["ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3, d["roofline"]["kernel_ms_alone_after_l2_flush"]*1e3))
PY
done
B200NB_PDL=1 timeout 600 python bench.py --no-cpu --workload water_192k > $O/bench_192k.json 2> $O/bench_192k.err
python - <<PY
import json
d=json.loads(open("$O/bench_192k.json").read().strip().splitlines()[-1]); print("192k step %.2f us  e2e %.2f us  kernel in step %.2f us"%(d["ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3))
PY
