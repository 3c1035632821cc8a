#!/bin/bash
# GPU pass ZX of round 2 (1 GPU): the one-GPU bench line once more on another box (the search step of ZW read 3.55 ms against 3.01 ms
# in ZO with identical search code: box-to-box variance of the host-synchronised parts, or not?)
mkdir -p gpurun_out/r2zx
O=gpurun_out/r2zx
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'search', d['search']['search_ms'], d['search']['search_ms_dynamic'])"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; lscpu | grep -i "model name\|^CPU(s)"
