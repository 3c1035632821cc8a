#!/bin/bash
mkdir -p gpurun_out
./profiles/tools/microbench > gpurun_out/microbench.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_domdec.py -x -q > gpurun_out/pytest_domdec.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_domdec.log
for n in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
cat gpurun_out/microbench.txt; tail -5 gpurun_out/pytest_domdec.log; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
