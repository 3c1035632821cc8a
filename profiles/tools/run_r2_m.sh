#!/bin/bash
# GPU pass M of round 2 (8 GPUs): strong scaling of the 1 M-atom box: 8 x-slabs, 2 x 2 x 2 domains (half shell through the
# peer-memory windows), 4 slabs, 2 x 2 x 1; weak scaling 24 k per GPU at N = 8
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $np --steps 50 --warmup 5 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d.get('parity'))" 2>&1 | tail -1)"
}
run bench_n8_slabs 8
run bench_n8_2x2x2 8 --dd-grid 2x2x2
run bench_n4_slabs 4
run bench_n4_2x2x1 4 --dd-grid 2x2x1
run bench_n8_weak24k 8 --workload water_24k --scaling weak
grep -h "Error\|error\|Traceback" $O/*.err | head -5
