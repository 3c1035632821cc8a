#!/bin/bash
# GPU pass ZR of round 2 (2 GPUs): the decomposed step's time-out flag mirrored into mapped host memory (b200nb_dd_status without a device
# call): DD suite, bench N = 2 (end-to-end figure)
mkdir -p gpurun_out/r2zr
O=gpurun_out/r2zr
timeout 1500 python -m pytest tests/test_gpu_domdec.py tests/test_gpu_domdec_ipc.py -q > $O/pytest_dd.txt 2>&1; tail -3 $O/pytest_dd.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"
python - <<'E'
import json
d=[json.loads(l) for l in open('gpurun_out/r2zr/bench_n2.json') if l.startswith('{')][0]
print('N=2 step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'], d['parity']['force_rel_rms_vs_single_domain'])
E
