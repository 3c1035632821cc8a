#!/bin/bash
# GPU pass M (8 GPUs): weak-scaling bench at N = 4 and 8 over peer-memory windows
mkdir -p gpurun_out/m
O=gpurun_out/m
nvidia-smi topo -m > $O/topo.txt 2>&1
for n in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 > $O/bench_n$n.json 2> $O/bench_n$n.err; cat $O/bench_n$n.json; grep -i "error\|Traceback" -A5 $O/bench_n$n.err | head -20
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 100 --warmup 10 --workload water_192k > $O/bench_n8_192k.json 2> $O/bench_n8_192k.err; cat $O/bench_n8_192k.json; grep -i "error\|Traceback" -A5 $O/bench_n8_192k.err | head -20
