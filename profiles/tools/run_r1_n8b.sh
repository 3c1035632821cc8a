#!/bin/bash
# GPU pass (8 GPUs): BASELINE configs[4] -- 12.3 M-atom LJ + reaction-field box, 1.536 M atoms per GPU (weak scaling); and 192 k atoms per GPU
N=8
mkdir -p gpurun_out/n8b
O=gpurun_out/n8b
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 3 --workload water_1.5M --eel rf > $O/bench_weak_water1.5M_rf.json 2> $O/bench_weak_water1.5M_rf.err; tail -c 600 $O/bench_weak_water1.5M_rf.json; tail -2 $O/bench_weak_water1.5M_rf.err
run --steps 50 --warmup 5 --workload water_192k > $O/bench_weak_water192k.json 2> $O/bench_weak_water192k.err; tail -c 600 $O/bench_weak_water192k.json; tail -2 $O/bench_weak_water192k.err
