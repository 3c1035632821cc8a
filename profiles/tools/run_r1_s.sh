#!/bin/bash
# GPU pass S (1 GPU): e2e A/B: host buffers accessed by the kernels (zero-copy) vs copy-engine staging; default entry length 24
mkdir -p gpurun_out/s
O=gpurun_out/s
for dma in 0 1 0 1; do
for wl in water_24k water_192k; do
B200NB_HOST_DMA=$dma timeout 600 python bench.py --no-cpu --workload $wl > $O/bench_dma${dma}_$wl.json 2> $O/bench_dma${dma}_$wl.err
python - <<PY
import json
d=json.loads(open("$O/bench_dma${dma}_$wl.json").read().strip().splitlines()[-1]); print("$wl HOST_DMA=$dma step %.2f us  e2e %.2f us  kernel in step %.2f us"%(d["ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3))
PY
done; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
B200NB_HOST_DMA=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
