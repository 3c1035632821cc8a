#!/bin/bash
# GPU pass K (2 GPUs): decomposed step with the halo chain on a high-priority non-local stream
mkdir -p gpurun_out/k
O=gpurun_out/k
timeout 600 python -m pytest tests/test_gpu_domdec.py -x -q > $O/pytest_domdec.log 2>&1; tail -2 $O/pytest_domdec.log
for wl in water_24k water_192k; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --workload $wl > $O/bench_n2_$wl.json 2> $O/bench_n2_$wl.err; cat $O/bench_n2_$wl.json; grep -i "error\|Traceback" -A5 $O/bench_n2_$wl.err | head -20
done
