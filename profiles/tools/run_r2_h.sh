#!/bin/bash
# GPU pass H of round 2 (2 GPUs): bench.py --gpus 2 on the new default (water_1M, strong) with the in-bench parity check
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"; tail -3 $O/bench_n2.err; cat $O/bench_n2.json | head -c 2500
