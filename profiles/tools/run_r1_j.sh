#!/bin/bash
# GPU pass J (2 GPUs): decomposed bench over peer-memory windows vs N=1
mkdir -p gpurun_out/j
O=gpurun_out/j
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; cat $O/bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > $O/bench_n2.json 2> $O/bench_n2.err; cat $O/bench_n2.json; tail -5 $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --workload water_192k > $O/bench_n2_192k.json 2> $O/bench_n2_192k.err; cat $O/bench_n2_192k.json; tail -3 $O/bench_n2_192k.err
