#!/usr/bin/env python
"""Times the force kernel alone (CUDA events, L2 flushed) for one build of the library.
usage: [B200NB_LIBRARY=scratch/lib_x.so] kernel_sweep.py <workload> <max_tiles,...> [ewald|rf]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import gmxapi_b200 as g

wl = sys.argv[1] if len(sys.argv) > 1 else "water_24k"
mts = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "16").split(",")]
eel = sys.argv[3] if len(sys.argv) > 3 else "ewald"
FLUSH = os.environ.get("NOFLUSH", "0") != "1"  # NOFLUSH=1: L2 stays warm between the timed launches
s = g.systems.named(wl)
for mt in mts:
    opt = g.NBKernelOptions(pairlistCutoff=0.9, coulombType=g.CoulombType.Pme if eel == "ewald" else g.CoulombType.ReactionField,
                            epsilonRf=0.0, maxTilesPerEntry=mt)
    fc = g.ForceCalculator(g.SimulationState.from_system(s), opt)
    st = fc.nb.stats()
    x = torch.from_numpy(s.x).cuda()
    f = torch.zeros_like(x)
    torch.cuda.synchronize()
    r = [fc.nb.time_step(x.data_ptr(), f.data_ptr(), 0, 3, 30, FLUSH) for _ in range(3)]
    step, ms = min(a for a, b in r), min(b for a, b in r)
    ms_cold = fc.nb.time_force_kernel(-1, 0, 3, 30, FLUSH)
    _, ms_vf = fc.nb.time_step(x.data_ptr(), f.data_ptr(), 3, 3, 30, FLUSH)
    print(("" if FLUSH else "[L2 warm] ") + "%s lib=%s max_tiles=%d entries=%d packed_tiles=%d: step %.2f us, force in step %.2f us (alone, cold L2 %.2f; VF %.2f), "
          "%.1f cycles/tile/SMSP" % (wl, os.path.basename(os.environ.get("B200NB_LIBRARY", "default")), mt, st["nentries"],
                                     st["ntiles_packed"], step * 1e3, ms * 1e3, ms_cold * 1e3, ms_vf * 1e3,
                                     ms * 1e-3 * 1.965e9 * 148 * 4 / st["ntiles_packed"]), flush=True)
    fc.nb.close()
