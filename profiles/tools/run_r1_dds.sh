#!/bin/bash
# GPU pass (N GPUs): decomposed step, two-stream schedule against the serial one (halo chain first on one stream)
N=$1
mkdir -p gpurun_out/dds$N
O=gpurun_out/dds$N
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N "$@"; }
for ser in 0 1; do
  for wl in water_24k water_96k; do
    B200NB_DD_SERIAL=$ser run --steps 200 --warmup 10 --workload $wl > $O/bench_serial${ser}_$wl.json 2> $O/bench_serial${ser}_$wl.err
    python - <<PY
import json
d=json.loads(open("$O/bench_serial${ser}_$wl.json").read().strip().splitlines()[-1]); print("$wl N=$N serial=$ser step %.2f us  e2e %.2f us  kernels %.2f us"%(d["ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3))
PY
  done
done
B200NB_DD_SERIAL=1 timeout 600 python -m pytest tests/test_gpu_domdec.py -x -q 2>&1 | tail -2
