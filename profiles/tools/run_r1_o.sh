#!/bin/bash
# GPU pass O (2 GPUs): A/B of graph replay and non-local priority for the decomposed step
mkdir -p gpurun_out/o
O=gpurun_out/o
run() { # name, workload, env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --workload $wl > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/$name.json").read().strip().splitlines()[-1]); print("$name", "step %.1f us  e2e %.1f us  kernel %.1f us"%(d["ms_per_step"]*1e3, d["e2e"]["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3))
except Exception as e: print("$name failed", e)
PY
}
for wl in water_24k water_192k; do
run graph_prio_$wl $wl B200NB_GRAPHS=1 B200NB_DD_PRIO=1
run direct_prio_$wl $wl B200NB_GRAPHS=0 B200NB_DD_PRIO=1
done
