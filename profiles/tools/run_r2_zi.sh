#!/bin/bash
# GPU pass ZI of round 2 (8 GPUs): strong scaling of the 1 M-atom box with the halo push on its own graph branch: N = 8 slabs and
# 2 x 2 x 2, N = 4 slabs, N = 2 slabs, N = 1 on the same box
mkdir -p gpurun_out/r2zi
O=gpurun_out/r2zi
run() { # name nproc args...
  name=$1; np=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $np --steps 50 --warmup 5 "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/$name.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d.get('parity'))" 2>&1 | tail -1)"
}
run bench_n8_slabs 8
run bench_n8_2x2x2 8 --dd-grid 2x2x2
run bench_n4_slabs 4
run bench_n2_slabs 2
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-sustained --no-search --secondary "" > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 rc=$? $(python -c "import json; d=json.load(open('$O/bench_n1.json')); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])")"
grep -h "Error\|error\|Traceback" $O/*.err | head -5
