#!/usr/bin/env python
"""Writes gmxapi_b200/data/ref_water_1000.npz: the base tile of the reference's own benchmark system -- coordinates1000, 1000
SPC/E molecules equilibrated at 300 K / 1 bar in a 3.10736 nm box (src/gromacs/nbnxm/benchmark/bench_coords.h:47-49) -- read out
of the reference through oracle/_ref (gmxref_bench_coordinates1000).  gmxapi_b200.systems.ref_water_box tiles it the way
BenchmarkSystem does (bench_system.cpp:90-151).  Run here; the file travels with the repo."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gmxref

L = gmxref.lib()
L.gmxref_bench_coordinates1000.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
edge = C.c_float()
n = L.gmxref_bench_coordinates1000(None, 0, C.byref(edge))
x = np.zeros((n, 3), np.float32)
assert L.gmxref_bench_coordinates1000(x.ctypes.data_as(C.c_void_p), n, C.byref(edge)) == n
out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "gmxapi_b200", "data", "ref_water_1000.npz")
np.savez_compressed(out, x=x, box_edge=np.float32(edge.value))
print("wrote", out, x.shape, edge.value)
