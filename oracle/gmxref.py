"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libgmxref_nbnxm.so.

The .so is the UNMODIFIED reference nbnxm CPU path (grid, pair search, plain-C / SIMD /
GPU-emulation kernels) compiled by oracle/build_ref.sh from /root/reference plus the C-ABI
harness oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgmxref_nbnxm.so")

KERNEL_PLAINC, KERNEL_SIMD_4XN, KERNEL_SIMD_2XNN, KERNEL_GPUREF = 0, 1, 2, 3
EEL_CUT, EEL_RF, EEL_EWALD_ANA, EEL_EWALD_TAB = 0, 1, 2, 3


class _System(C.Structure):
    _fields_ = [("natoms", C.c_int), ("x", C.c_void_p), ("box", C.c_float * 3), ("ntypes", C.c_int),
                ("nbfp", C.c_void_p), ("type", C.c_void_p), ("q", C.c_void_p),
                ("excl_off", C.c_void_p), ("excl_idx", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("rc", C.c_float), ("rlist", C.c_float), ("rlist_inner", C.c_float),
                ("nstlist_prune", C.c_int), ("eeltype", C.c_int), ("epsfac", C.c_float),
                ("k_rf", C.c_float), ("c_rf", C.c_float), ("ewaldcoeff", C.c_float),
                ("sh_ewald", C.c_float), ("disp_cpot", C.c_float), ("rep_cpot", C.c_float),
                ("kernel", C.c_int), ("comb_rule", C.c_int), ("nthreads", C.c_int),
                ("exact_atom_flags", C.c_int), ("put_in_box", C.c_int), ("min_ilist_count", C.c_int),
                ("rvdw", C.c_float), ("vdw_modifier", C.c_int), ("rvdw_switch", C.c_float),
                ("disp_c2", C.c_float), ("disp_c3", C.c_float), ("rep_c2", C.c_float), ("rep_c3", C.c_float),
                ("sw_c3", C.c_float), ("sw_c4", C.c_float), ("sw_c5", C.c_float),
                ("ljpme", C.c_int), ("ewaldcoeff_lj", C.c_float), ("sh_lj_ewald", C.c_float),
                ("box_offdiag", C.c_float * 3), ("perturbed", C.c_void_p)]


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        # bind OpenMP threads before libgomp initialises (unbound threads are >3x slower)
        os.environ.setdefault("OMP_PROC_BIND", "spread")
        os.environ.setdefault("OMP_PLACES", "cores")
        L = C.CDLL(LIB_PATH)
        L.gmxref_create.restype = C.c_void_p
        L.gmxref_create.argtypes = [C.POINTER(_System), C.POINTER(_Params)]
        L.gmxref_destroy.argtypes = [C.c_void_p]
        L.gmxref_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gmxref_time_kernel.restype = C.c_double
        L.gmxref_time_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.gmxref_time_step.restype = C.c_double
        L.gmxref_time_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.gmxref_grid_order.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.gmxref_grid_dims.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.gmxref_list_stats.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.gmxref_pair_set.restype = C.c_longlong
        L.gmxref_pair_set.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_longlong]
        L.gmxref_regrid_research.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gmxref_setup_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gmxref_ewald_coeff.restype = C.c_float
        L.gmxref_ewald_coeff.argtypes = [C.c_float, C.c_float]
        L.gmxref_simd_rsq.restype = C.c_float
        L.gmxref_simd_rsq.argtypes = [C.c_float] * 6
        L.gmxref_gpu_list.argtypes = [C.c_void_p] * 10
        L.gmxref_grid_forces.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.gmxref_ewald_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def ewald_coeff(rc, rtol=1e-5):
    return float(lib().gmxref_ewald_coeff(rc, rtol))


def default_simd_kernel():
    return int(lib().gmxref_default_simd_kernel())


class RefNbnxm:
    """One reference nonbonded_verlet_t instance: gridded, searched, ready to compute."""

    def __init__(self, x, box, types, q, nbfp, excl_off, excl_idx, rc, rlist=None, eeltype=EEL_CUT,
                 epsfac=138.935458, k_rf=0.0, c_rf=0.0, ewaldcoeff=0.0, sh_ewald=0.0,
                 disp_cpot=None, rep_cpot=None, kernel=None, comb_rule=0, nthreads=1,
                 exact_atom_flags=0, put_in_box=0, rlist_inner=0.0, min_ilist_count=0,
                 rvdw=0.0, vdw_modifier=0, rvdw_switch=0.0, modifier_constants=None, ljpme=0, ewaldcoeff_lj=0.0,
                 sh_lj_ewald=0.0, box_offdiag=(0.0, 0.0, 0.0), perturbed=None):
        """box: the diagonal of the box matrix; box_offdiag: box[YY][XX], box[ZZ][XX], box[ZZ][YY] of a triclinic cell"""
        L = lib()
        pert = None if perturbed is None else np.ascontiguousarray(perturbed, dtype=np.uint8)  # read during gmxref_create only
        self.n = int(len(types))
        self._x = np.ascontiguousarray(x, dtype=np.float32).reshape(self.n, 3)
        self._types = np.ascontiguousarray(types, dtype=np.int32)
        self._q = np.ascontiguousarray(q, dtype=np.float32)
        self._nbfp = np.ascontiguousarray(nbfp, dtype=np.float32).ravel()
        ntypes = int(round((self._nbfp.size // 2) ** 0.5))
        self._eo = np.ascontiguousarray(excl_off, dtype=np.int32)
        self._ei = np.ascontiguousarray(excl_idx, dtype=np.int32)
        s = _System(self.n, self._x.ctypes.data, (C.c_float * 3)(*[float(b) for b in box]), ntypes,
                    self._nbfp.ctypes.data, self._types.ctypes.data, self._q.ctypes.data,
                    self._eo.ctypes.data, self._ei.ctypes.data)
        if kernel is None:
            kernel = default_simd_kernel()
        # modifier_constants: dict from oracle.vdw_modifier_constants (the values init_interaction_const would store)
        k = modifier_constants or {}
        rv = rvdw if rvdw > 0 else rc
        if disp_cpot is None:
            disp_cpot = k.get("disp_cpot", -1.0 / rv ** 6)
        if rep_cpot is None:
            rep_cpot = k.get("rep_cpot", -1.0 / rv ** 12)
        if vdw_modifier != 0 and not k:
            raise ValueError("vdw_modifier needs modifier_constants")
        p = _Params(rc, rlist if rlist else rc, rlist_inner, 0, eeltype, epsfac, k_rf, c_rf, ewaldcoeff,
                    sh_ewald, disp_cpot, rep_cpot, kernel, comb_rule, nthreads, exact_atom_flags,
                    put_in_box, min_ilist_count, rvdw, vdw_modifier, rvdw_switch,
                    k.get("disp_c2", 0.0), k.get("disp_c3", 0.0), k.get("rep_c2", 0.0), k.get("rep_c3", 0.0),
                    k.get("sw_c3", 0.0), k.get("sw_c4", 0.0), k.get("sw_c5", 0.0), ljpme, ewaldcoeff_lj, sh_lj_ewald,
                    (C.c_float * 3)(*[float(v) for v in box_offdiag]),
                    None if perturbed is None else pert.ctypes.data)
        self.rc = rc
        self.h = L.gmxref_create(C.byref(s), C.byref(p))
        if not self.h:
            raise RuntimeError("gmxref_create failed (kernel type unavailable in this build?)")

    def close(self):
        if self.h:
            lib().gmxref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, x=None, energy=True, virial=True):
        f = np.zeros((self.n, 3), np.float32)
        fs = np.zeros((45, 3), np.float32)
        e = np.zeros(2, np.float32)
        xp = None
        if x is not None:
            xx = np.ascontiguousarray(x, dtype=np.float32)
            xp = xx.ctypes.data
        lib().gmxref_compute(self.h, xp, int(energy), int(virial), f.ctypes.data, fs.ctypes.data, e.ctypes.data)
        return f, fs, float(e[0]), float(e[1])

    def time_kernel(self, energy=False, nwarm=1, niter=5):
        return float(lib().gmxref_time_kernel(self.h, int(energy), nwarm, niter))

    def time_step(self, energy=False, nwarm=1, niter=5):
        return float(lib().gmxref_time_step(self.h, int(energy), nwarm, niter))

    def regrid_research(self):
        tg, ts = C.c_double(), C.c_double()
        lib().gmxref_regrid_research(self.h, C.byref(tg), C.byref(ts))
        return tg.value, ts.value

    def setup_times(self):
        tg, ts = C.c_double(), C.c_double()
        lib().gmxref_setup_times(self.h, C.byref(tg), C.byref(ts))
        return tg.value, ts.value

    def grid_order(self):
        cap = self.n * 2 + 4096
        out = np.zeros(cap, np.int32)
        n = lib().gmxref_grid_order(self.h, out.ctypes.data, cap)
        return out[:n].copy()

    def grid_dims(self):
        ncx, ncy, npad = C.c_int(), C.c_int(), C.c_int()
        cx, cy = C.c_float(), C.c_float()
        lib().gmxref_grid_dims(self.h, C.byref(ncx), C.byref(ncy), C.byref(cx), C.byref(cy), C.byref(npad))
        return ncx.value, ncy.value, cx.value, cy.value, npad.value

    def list_stats(self):
        a, b = C.c_longlong(), C.c_longlong()
        ci, cj = C.c_int(), C.c_int()
        lib().gmxref_list_stats(self.h, C.byref(a), C.byref(b), C.byref(ci), C.byref(cj))
        return dict(cluster_pairs=a.value, atom_pairs_computed=b.value, na_ci=ci.value, na_cj=cj.value)

    def gpu_list(self):
        """The reference-built 8x8x8 list and grid-ordered atom data (GPUREF instances): dict(sci (n,4) int32, cj4 (n,8) int32,
        excl (n,32) uint32, xq (nslots,4) float32, type (nslots,) int32): the arguments of gpu_init_pairlist / gpu_init_atomdata."""
        L = lib()
        n = [C.c_int() for _ in range(4)]
        if L.gmxref_gpu_list(self.h, *[C.byref(v) for v in n], None, None, None, None, None):
            raise RuntimeError("gpu_list needs a GPUREF_8X8X8 instance")
        nsci, ncj4, nexcl, nslots = [v.value for v in n]
        out = dict(sci=np.zeros((nsci, 4), np.int32), cj4=np.zeros((ncj4, 8), np.int32), excl=np.zeros((nexcl, 32), np.uint32),
                   xq=np.zeros((nslots, 4), np.float32), type=np.zeros(nslots, np.int32))
        rc = L.gmxref_gpu_list(self.h, *[C.byref(v) for v in n], *[out[k].ctypes.data_as(C.c_void_p) for k in ("sci", "cj4", "excl", "xq", "type")])
        if rc:
            raise RuntimeError("gmxref_gpu_list failed: %d" % rc)
        return out

    def ewald_table(self):
        """(tableF float32 array, scale): the reference's tabulated Ewald force correction of this instance."""
        sc = C.c_float()
        n = lib().gmxref_ewald_table(self.h, None, 0, C.byref(sc))
        t = np.zeros(n, np.float32)
        lib().gmxref_ewald_table(self.h, t.ctypes.data_as(C.c_void_p), n, C.byref(sc))
        return t, float(sc.value)

    def grid_forces(self):
        """nbat->out[0].f of the last compute(), grid order (nslots, 3)."""
        n = self.grid_dims()[4]
        f = np.zeros((n, 3), np.float32)
        if lib().gmxref_grid_forces(self.h, f.ctypes.data_as(C.c_void_p), n) != n:
            raise RuntimeError("gmxref_grid_forces failed")
        return f

    def pair_count(self, rc=None):
        return int(lib().gmxref_pair_set(self.h, rc or self.rc, None, 0))

    def pair_set(self, rc=None):
        """(npairs, 3) int32 array of (i_atom [shifted], j_atom, shift index)."""
        n = self.pair_count(rc)
        out = np.zeros((max(n, 1), 3), np.int32)
        lib().gmxref_pair_set(self.h, rc or self.rc, out.ctypes.data, n)
        return out[:n]


class _FepParams(C.Structure):
    _fields_ = [("rc", C.c_float), ("epsfac", C.c_float), ("k_rf", C.c_float), ("c_rf", C.c_float), ("disp_cpot", C.c_float),
                ("rep_cpot", C.c_float), ("lambda_coul", C.c_float), ("lambda_vdw", C.c_float), ("sc_alpha", C.c_float),
                ("sc_power", C.c_int), ("sc_sigma", C.c_float), ("sc_sigma_min", C.c_float), ("sc_coul", C.c_int),
                ("ewaldcoeff", C.c_float), ("sh_ewald", C.c_float), ("rvdw_switch", C.c_float),
                ("ljpme", C.c_int), ("ewaldcoeff_lj", C.c_float), ("sh_lj_ewald", C.c_float), ("rvdw", C.c_float)]


def fep_list(ref):
    """The perturbed pair lists the reference's own search built for a Reference(..., perturbed=flags): nbnxm/pairlist.cpp
    make_fep_list through PairlistSet::fepLists(), concatenated over the search threads.  Returns iinr, shift, jindex, jjnr,
    excl_fep."""
    L = lib()
    nri, nrj = C.c_int(0), C.c_int(0)
    vp = C.c_void_p
    L.gmxref_fep_list.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.gmxref_fep_list(ref.h, C.addressof(nri), C.addressof(nrj), 0, 0, None, None, None, None, None)
    ii, sh, ji = np.zeros(nri.value, np.int32), np.zeros(nri.value, np.int32), np.zeros(nri.value + 1, np.int32)
    jj, ex = np.zeros(max(nrj.value, 1), np.int32), np.zeros(max(nrj.value, 1), np.int8)
    rc = L.gmxref_fep_list(ref.h, C.addressof(nri), C.addressof(nrj), len(ii), len(jj), ii.ctypes.data, sh.ctypes.data, ji.ctypes.data,
                           jj.ctypes.data, ex.ctypes.data)
    if rc != 0:
        raise RuntimeError("gmxref_fep_list failed")
    return ii, sh, ji, jj[:nrj.value], ex[:nrj.value]


def fep_kernel(x, shift_vec, nbfp, typeA, typeB, qA, qB, iinr, shift, jindex, jjnr, excl_fep, rc, lambda_coul, lambda_vdw,
               epsfac=138.935458, k_rf=0.0, c_rf=0.0, sc_alpha=0.5, sc_power=1, sc_sigma=0.3, sc_sigma_min=0.3, sc_coul=False,
               ewaldcoeff=0.0, sh_ewald=0.0, rvdw_switch=0.0, ljpme=0, ewaldcoeff_lj=0.0, sh_lj_ewald=0.0, rvdw=0.0):
    """The reference's gmx_nb_free_energy_kernel (gmxlib/nonbonded/nb_free_energy.cpp) on a perturbed pair list in t_nblist form.
    Returns f[n,3], fshift[45,3], (Vc, Vv, dvdl_coul, dvdl_vdw)."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3)
    n = x.shape[0]
    sv = np.ascontiguousarray(shift_vec, dtype=np.float32).reshape(45, 3)
    nb = np.ascontiguousarray(nbfp, dtype=np.float32).ravel()
    ntypes = int(round((nb.size // 2) ** 0.5))
    arr = lambda a, t: np.ascontiguousarray(a, dtype=t)
    tA, tB, cA, cB = arr(typeA, np.int32), arr(typeB, np.int32), arr(qA, np.float32), arr(qB, np.float32)
    ii, sh, ji, jj = arr(iinr, np.int32), arr(shift, np.int32), arr(jindex, np.int32), arr(jjnr, np.int32)
    ex = arr(excl_fep, np.int8)
    cpot6, cpot12 = (0.0, 0.0) if rvdw_switch > 0 else (-1.0 / (rvdw or rc) ** 6, -1.0 / (rvdw or rc) ** 12)  # rvdw_switch > 0: LJ potential switch
    p = _FepParams(rc, epsfac, k_rf, c_rf, cpot6, cpot12, lambda_coul, lambda_vdw, sc_alpha, sc_power, sc_sigma,
                   sc_sigma_min, int(bool(sc_coul)), ewaldcoeff, sh_ewald, rvdw_switch, int(ljpme), ewaldcoeff_lj, sh_lj_ewald, rvdw)
    f = np.zeros((n, 3), np.float32)
    fs = np.zeros((45, 3), np.float32)
    out = np.zeros(4, np.float32)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    L.gmxref_fep_kernel.restype = C.c_int
    L.gmxref_fep_kernel(C.c_int(n), vp(x), vp(sv), C.c_int(ntypes), vp(nb), vp(tA), vp(tB), vp(cA), vp(cB), C.c_int(len(ii)), vp(ii), vp(sh),
                        vp(ji), vp(jj), vp(ex), C.byref(p), vp(f), vp(fs), vp(out))
    return f, fs, tuple(float(v) for v in out)


def bonded(kind, iatoms, params6, x, q, box_matrix, epsfac_fudge=138.935458 * 0.5, virial_energy=True):
    """The reference's CPU functions for the listed interactions its GPU bonded module covers (gmxref_bonded: bonded.cpp
    calculateSimpleBond, pairs.cpp do_pairs).  Arguments as oracle.bonded.  Returns f[n,3], fshift[45,3] (float32), energy; for
    lj14 forces only (the reference's analytical path has no energy / virial: its general path uses spline tables)."""
    from .oracle import BONDED_KINDS, BONDED_NRAL
    k = BONDED_KINDS.index(kind) if isinstance(kind, str) else int(kind)
    ia = np.ascontiguousarray(iatoms, dtype=np.int32).reshape(-1, BONDED_NRAL[k] + 1)
    p6 = np.ascontiguousarray(params6, dtype=np.float32).reshape(-1, 6)
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3)
    n = x.shape[0]
    q = np.ascontiguousarray(q, dtype=np.float32)
    b = np.ascontiguousarray(box_matrix, dtype=np.float32).reshape(9)
    f = np.zeros((n, 3), np.float32)
    fs = np.zeros((45, 3), np.float32)
    e = np.zeros(2, np.float64)
    L = lib()
    L.gmxref_bonded.restype = C.c_int
    vp = C.c_void_p
    L.gmxref_bonded.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, vp, vp, vp]
    rc = L.gmxref_bonded(k, len(ia), ia.ctypes.data, len(p6), p6.ctypes.data, n, x.ctypes.data, q.ctypes.data, b.ctypes.data,
                         float(epsfac_fudge), int(bool(virial_energy)), f.ctypes.data, fs.ctypes.data, e.ctypes.data)
    if rc != 0:
        raise ValueError("gmxref_bonded failed (%d)" % rc)
    return f, fs, float(e[0])
