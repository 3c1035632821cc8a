#!/bin/bash
# GPU pass ZC of round 2 (2 GPUs): halo x push as its own kernel at the head of the non-local stream; DD tests (threads on one GPU and
# two processes over CUDA IPC), bench N = 2 against N = 1 on the same box; pair-search step with 6 / 8 / 10 resident search CTAs per SM
mkdir -p gpurun_out/r2zc
O=gpurun_out/r2zc
timeout 1500 python -m pytest tests/test_gpu_domdec.py tests/test_gpu_domdec_ipc.py -q > $O/pytest_dd.txt 2>&1; tail -4 $O/pytest_dd.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"; head -c 300 $O/bench_n2.json; echo
for lib in default; do
  L=""; [ $lib != default ] && L=scratch/lib_$lib.so
  B200NB_LIBRARY=$L timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-sustained --secondary "" > $O/bench_$lib.json 2> $O/bench_$lib.err
  python - <<E
import json
d=json.load(open('$O/bench_$lib.json'))
print('$lib', 'step', round(d['ms_per_step'],4), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
E
done
python - <<'E'
import json
d=[json.loads(l) for l in open('gpurun_out/r2zc/bench_n2.json') if l.startswith('{')][0]
print('N=2 step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'], d['parity']['force_rel_rms_vs_single_domain'])
E
