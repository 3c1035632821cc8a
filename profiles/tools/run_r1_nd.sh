#!/bin/bash
# GPU pass (4 GPUs): 2 x 2 x 1 half-shell decomposition over NCCL against 4 x-slabs with peer-memory windows, 1 M-atom box
N=4
mkdir -p gpurun_out/nd4
O=gpurun_out/nd4
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N "$@"; }
run --steps 30 --warmup 3 --workload water_1M --scaling strong --dd-grid 2x2x1 > $O/bench_strong_water1M_2x2x1.json 2> $O/bench_strong_water1M_2x2x1.err
tail -c 500 $O/bench_strong_water1M_2x2x1.json; tail -4 $O/bench_strong_water1M_2x2x1.err
