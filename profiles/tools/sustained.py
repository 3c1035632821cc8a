#!/usr/bin/env python
"""sustained.py <workload> <seconds>: back-to-back device-resident steps (graph replays, no L2 flush) for `seconds`, NVML sampled
every 2 ms: median SM clock, throttle reasons, power, step time at that clock.  (bench.py prints the same block.)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import gmxapi_b200 as g
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "water_24k"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
s = g.systems.named(wl)
opt = g.NBKernelOptions(pairlistCutoff=0.9, coulombType=g.CoulombType.Pme)
fc = g.ForceCalculator(g.SimulationState.from_system(s), opt)
print(bench.sustained_block(fc.nb, s, fc.nb.pair_count(0.9), secs))
