#!/bin/bash
# GPU pass (N GPUs, $1 = N): decomposed step with the ring kernel: default weak-scaling line, strong scaling of the 1 M-atom box; DD parity tests
N=$1
mkdir -p gpurun_out/n$N
O=gpurun_out/n$N
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 200 --warmup 10 > $O/bench_weak_water24k.json 2> $O/bench_weak_water24k.err; tail -c 700 $O/bench_weak_water24k.json; tail -2 $O/bench_weak_water24k.err
run --steps 50 --warmup 5 --workload water_1M --scaling strong > $O/bench_strong_water1M.json 2> $O/bench_strong_water1M.err; tail -c 700 $O/bench_strong_water1M.json; tail -2 $O/bench_strong_water1M.err
if [ "$2" = "tests" ]; then timeout 900 python -m pytest tests/test_gpu_domdec.py -x -q 2>&1 | tail -3 | tee $O/pytest_dd.txt; fi
if [ "$2" = "rf12m" ]; then run --steps 20 --warmup 3 --workload water_1.5M --eel rf > $O/bench_weak_water1.5M_rf.json 2> $O/bench_weak_water1.5M_rf.err; tail -c 700 $O/bench_weak_water1.5M_rf.json; tail -2 $O/bench_weak_water1.5M_rf.err; fi
