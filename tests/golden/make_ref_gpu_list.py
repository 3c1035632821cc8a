#!/usr/bin/env python
"""Generates tests/golden/ref_gpu_list_water_3k.npz: the REFERENCE-built 8x8x8 pair list (NbnxnPairlistGpu: sci, cj4, excl) and
grid-ordered atom data (nbat x / type, atom indices) of the 3000-atom water box, exactly as Nbnxm::gpu_init_pairlist /
gpu_init_atomdata receive them (nbnxm_gpu_data_mgmt.cpp:251-311), together with the reference's own outputs on that list
(nbnxn_kernel_gpu_ref, kernels_reference/kernel_gpu_ref.cpp, reaction field): grid-ordered forces, shift forces, energies.
Run here (needs oracle/_ref built from /root/reference by oracle/build_ref.sh); the fixture travels to the GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gmxapi_b200 as g
from oracle import gmxref

RC = 0.9
s = g.systems.named("water_3k")
k_rf, c_rf = g.systems.rf_constants(RC, eps_rf=1.0)
r = gmxref.RefNbnxm(s.x, s.box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx, rc=RC, eeltype=gmxref.EEL_RF, k_rf=k_rf, c_rf=c_rf,
                    kernel=gmxref.KERNEL_GPUREF, nthreads=1)
f, fshift, elj, eel = r.compute(energy=True, virial=True)
L = r.gpu_list()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_gpu_list_water_3k.npz")
np.savez_compressed(out, sci=L["sci"], cj4=L["cj4"], excl=L["excl"], xq=L["xq"], type=L["type"], atom_index=r.grid_order(),
                    f_grid=r.grid_forces(), f=f, fshift=fshift, energies=np.array([elj, eel], np.float64), rc=RC, k_rf=k_rf, c_rf=c_rf,
                    box=np.asarray(s.box, np.float32), npairs=r.pair_count())
print("wrote", out, {k: v.shape for k, v in L.items()}, "pairs", r.pair_count(), "E", elj, eel)
