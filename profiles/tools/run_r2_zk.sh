#!/bin/bash
# GPU pass ZK of round 2 (8 GPUs): BASELINE configs[4], the 12.3 M-atom LJ + reaction-field box as a weak-scaling run (1.536 M atoms
# per GPU) with the end-of-round kernel, and the 1.536 M-atom box on one GPU of the same box
mkdir -p gpurun_out/r2zk
O=gpurun_out/r2zk
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 --workload water_1.5M --eel rf --scaling weak --no-parity > $O/bench_n8_weak_water1.5M_rf.json 2> $O/bench_n8_weak.err; echo "n8 weak rc=$?"; tail -2 $O/bench_n8_weak.err
timeout 300 python bench.py --steps 20 --warmup 3 --workload water_1.5M --eel rf --no-cpu --no-sustained --no-search --secondary "" > $O/bench_n1_water1.5M_rf.json 2> $O/bench_n1.err; echo "n1 rc=$?"
python - <<'E'
import json
for n in ("bench_n8_weak_water1.5M_rf","bench_n1_water1.5M_rf"):
    try:
        d=[json.loads(l) for l in open('gpurun_out/r2zk/%s.json'%n) if l.startswith('{')][0]
        print(n, d['config'].get('atoms'), 'step', d['ms_per_step'], 'value %.4g'%d['value'], 'kernel', d['roofline']['kernel_ms'], 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e']['ms_per_step'])
    except Exception as e: print(n, 'failed', e)
E
