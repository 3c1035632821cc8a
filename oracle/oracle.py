"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the plain-C oracle (oracle/nbnxm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (gmxapi_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libnbnxm_oracle.so")
EEL_CUT, EEL_RF, EEL_EWALD = 0, 1, 2
CENTRAL, SHIFTS = 22, 45


class _Params(C.Structure):
    _fields_ = [("rc", C.c_float), ("eeltype", C.c_int), ("epsfac", C.c_float), ("k_rf", C.c_float),
                ("c_rf", C.c_float), ("beta", C.c_float), ("sh_ewald", C.c_float),
                ("disp_cpot", C.c_float), ("rep_cpot", C.c_float), ("ntypes", C.c_int), ("nbfp", C.c_void_p),
                ("rvdw", C.c_float), ("vdw_modifier", C.c_int), ("rvdw_switch", C.c_float),
                ("disp_c2", C.c_float), ("disp_c3", C.c_float), ("rep_c2", C.c_float), ("rep_c3", C.c_float),
                ("sw_c3", C.c_float), ("sw_c4", C.c_float), ("sw_c5", C.c_float),
                ("ljpme", C.c_int), ("ewaldcoeff_lj", C.c_float), ("sh_lj_ewald", C.c_float)]


VDW_POTSHIFT, VDW_FORCESWITCH, VDW_POTSWITCH = 0, 1, 2


LJPME_NONE, LJPME_GEOM, LJPME_LB = 0, 1, 2


def lj_ewald_shift(ewaldcoeff_lj, rvdw):
    """interaction_const_t::sh_lj_ewald with a potential-shift modifier (mdlib/forcerec.cpp:709-713)."""
    import math
    crc2 = (ewaldcoeff_lj * rvdw) ** 2
    return (math.exp(-crc2) * (1 + crc2 + 0.5 * crc2 * crc2) - 1) / rvdw ** 6


def vdw_modifier_constants(modifier, rvdw, rvdw_switch):
    """interaction_const_t's LJ modifier constants as init_interaction_const makes them (mdlib/forcerec.cpp:850-874):
    dict(disp_cpot, rep_cpot, disp_c2, disp_c3, rep_c2, rep_c3, sw_c3, sw_c4, sw_c5)."""
    d = dict(disp_cpot=0.0, rep_cpot=0.0, disp_c2=0.0, disp_c3=0.0, rep_c2=0.0, rep_c3=0.0, sw_c3=0.0, sw_c4=0.0, sw_c5=0.0)
    if modifier == VDW_POTSHIFT:
        d["disp_cpot"], d["rep_cpot"] = -1.0 / rvdw ** 6, -1.0 / rvdw ** 12
    elif modifier == VDW_FORCESWITCH:
        o = (C.c_float * 3)()
        lib().orc_force_switch_constants(C.c_double(6.0), C.c_double(rvdw_switch), C.c_double(rvdw), o)
        d["disp_c2"], d["disp_c3"], d["disp_cpot"] = float(o[0]), float(o[1]), float(o[2])
        lib().orc_force_switch_constants(C.c_double(12.0), C.c_double(rvdw_switch), C.c_double(rvdw), o)
        d["rep_c2"], d["rep_c3"], d["rep_cpot"] = float(o[0]), float(o[1]), float(o[2])
    else:
        o = (C.c_float * 3)()
        lib().orc_potential_switch_constants(C.c_double(rvdw_switch), C.c_double(rvdw), o)
        d["sw_c3"], d["sw_c4"], d["sw_c5"] = float(o[0]), float(o[1]), float(o[2])
    return d


_lib = None


def build():
    src = os.path.join(_HERE, "nbnxm_oracle.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "_build/libnbnxm_oracle.so"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_rsq.restype = C.c_float
        L.orc_rsq.argtypes = [C.c_float] * 9
        L.orc_put_on_grid.restype = C.c_int
        L.orc_pair_set.restype = C.c_longlong
        L.orc_tile_list.restype = C.c_longlong
        L.orc_forces.restype = C.c_longlong
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def rsq(xi, shift, xj):
    return float(lib().orc_rsq(*[C.c_float(float(v)) for v in (*xi, *shift, *xj)]))


def set_triclinic(offdiag=None):
    """Triclinic mode of the oracle: the off-diagonal box elements box[YY][XX], box[ZZ][XX], box[ZZ][YY] (None / zeros:
    rectangular).  A mode, not an argument: every function keeps taking the box DIAGONAL.  Pair enumeration becomes brute force."""
    if offdiag is None:
        lib().orc_set_triclinic(C.c_void_p(0))
    else:
        lib().orc_set_triclinic((C.c_float * 3)(*[float(v) for v in offdiag]))


def shift_vectors(box):
    sv = np.zeros((SHIFTS, 3), np.float32)
    b = (C.c_float * 3)(*[float(v) for v in box])
    lib().orc_shift_vectors(b, _p(sv))
    return sv


def put_on_grid(x, box, density=None):
    """Returns dict(ncx, ncy, col_cell0, atom_index, slot_of_atom)."""
    x = _f32(x)
    n = x.shape[0]
    lower = (C.c_float * 3)(0, 0, 0)
    upper = (C.c_float * 3)(*[float(v) for v in box])
    if density is None:
        density = np.float32(n) / (np.float32(box[0]) * np.float32(box[1]) * np.float32(box[2]))
    cap = n + 64 * 70000 + 64
    colcap = 70002
    atom_index = np.zeros(cap, np.int32)
    col_cell0 = np.zeros(colcap, np.int32)
    slot = np.zeros(n, np.int32)
    ncx, ncy = C.c_int(), C.c_int()
    npad = lib().orc_put_on_grid(C.c_int(n), _p(x), lower, upper, C.c_float(float(density)), C.byref(ncx),
                                 C.byref(ncy), _p(col_cell0), C.c_int(colcap), _p(atom_index), C.c_int(cap),
                                 _p(slot))
    if npad < 0:
        raise RuntimeError("orc_put_on_grid: capacity exceeded")
    ncol = ncx.value * ncy.value
    return dict(ncx=ncx.value, ncy=ncy.value, col_cell0=col_cell0[:ncol + 1].copy(),
                atom_index=atom_index[:npad].copy(), slot_of_atom=slot, npad=npad)


def pair_set(x, box, r, excl_off=None, excl_idx=None):
    """(npairs,3) int32: (i [the shifted atom], j, shift index), unsorted."""
    x = _f32(x)
    n = x.shape[0]
    b = (C.c_float * 3)(*[float(v) for v in box])
    eo = _i32(excl_off) if excl_off is not None else None
    ei = _i32(excl_idx) if excl_idx is not None else None
    cnt = lib().orc_pair_set(C.c_int(n), _p(x), b, C.c_float(r), _p(eo), _p(ei), C.c_void_p(0), C.c_longlong(0))
    out = np.zeros((max(cnt, 1), 3), np.int32)
    lib().orc_pair_set(C.c_int(n), _p(x), b, C.c_float(r), _p(eo), _p(ei), _p(out), C.c_longlong(cnt))
    return out[:cnt]


def canonical_pairs(pairs):
    """Sort (i,j,shift) triples into a canonical order; CENTRAL pairs get i<j."""
    p = np.array(pairs, dtype=np.int64, copy=True).reshape(-1, 3)
    cen = p[:, 2] == CENTRAL
    lo = np.minimum(p[:, 0], p[:, 1])
    hi = np.maximum(p[:, 0], p[:, 1])
    p[cen, 0] = lo[cen]
    p[cen, 1] = hi[cen]
    key = (p[:, 0] << 34) | (p[:, 1] << 6) | p[:, 2]
    return np.sort(key)


def tile_list(x, box, rlist, slot_of_atom):
    x = _f32(x)
    n = x.shape[0]
    b = (C.c_float * 3)(*[float(v) for v in box])
    slot = _i32(slot_of_atom)
    cap = max(1 << 16, n * 24)
    while True:
        out = np.zeros((cap, 3), np.int32)
        m = lib().orc_tile_list(C.c_int(n), _p(x), b, C.c_float(rlist), _p(slot), _p(out), C.c_longlong(cap))
        if m >= 0:
            return out[:m]
        cap *= 4


def prune_tiles(tiles, atom_index, x, box, rlist_inner):
    tiles = _i32(tiles)
    x = _f32(x)
    ai = _i32(atom_index)
    b = (C.c_float * 3)(*[float(v) for v in box])
    keep = np.zeros(len(tiles), np.uint8)
    lib().orc_prune_tiles(C.c_longlong(len(tiles)), _p(tiles), _p(ai), _p(x), b, C.c_float(rlist_inner), _p(keep))
    return keep.astype(bool)


def forces(x, box, q, types, nbfp, rc, excl_off=None, excl_idx=None, eeltype=EEL_CUT, epsfac=138.935458,
           k_rf=0.0, c_rf=0.0, beta=0.0, sh_ewald=0.0, disp_cpot=None, rep_cpot=None, energy=True,
           rvdw=0.0, vdw_modifier=VDW_POTSHIFT, rvdw_switch=0.0, ljpme=0, ewaldcoeff_lj=0.0, sh_lj_ewald=None):
    """Returns (f[n,3] float64, fshift[45,3] float64, evdw, ecoul, npairs).  rvdw < rc: twin-range cut-off (Ewald only in
    the reference); vdw_modifier: VDW_POTSHIFT / VDW_FORCESWITCH / VDW_POTSWITCH from rvdw_switch to rvdw."""
    x = _f32(x)
    n = x.shape[0]
    q = _f32(q)
    types = _i32(types)
    nbfp = _f32(nbfp).ravel()
    ntypes = int(round((nbfp.size // 2) ** 0.5))
    rv = rvdw if rvdw > 0 else rc
    k = vdw_modifier_constants(vdw_modifier, rv, rvdw_switch)
    if disp_cpot is None:
        disp_cpot = k["disp_cpot"]
    if rep_cpot is None:
        rep_cpot = k["rep_cpot"]
    p = _Params(rc, eeltype, epsfac, k_rf, c_rf, beta, sh_ewald, disp_cpot, rep_cpot, ntypes, nbfp.ctypes.data,
                rvdw, vdw_modifier, rvdw_switch, k["disp_c2"], k["disp_c3"], k["rep_c2"], k["rep_c3"],
                k["sw_c3"], k["sw_c4"], k["sw_c5"], ljpme, ewaldcoeff_lj,
                lj_ewald_shift(ewaldcoeff_lj, rv) if (sh_lj_ewald is None and ljpme) else (sh_lj_ewald or 0.0))
    b = (C.c_float * 3)(*[float(v) for v in box])
    eo = _i32(excl_off) if excl_off is not None else None
    ei = _i32(excl_idx) if excl_idx is not None else None
    f = np.zeros((n, 3), np.float64)
    fs = np.zeros((SHIFTS, 3), np.float64)
    e = np.zeros(2, np.float64)
    npairs = lib().orc_forces(C.c_int(n), _p(x), b, _p(q), _p(types), _p(eo), _p(ei), C.byref(p),
                              C.c_int(int(energy)), _p(f), _p(fs), _p(e))
    return f, fs, float(e[0]), float(e[1]), int(npairs)


def virial_from_fshift(box, fshift):
    b = (C.c_float * 3)(*[float(v) for v in box])
    fs = np.ascontiguousarray(fshift, dtype=np.float64)
    vir = np.zeros(9, np.float64)
    lib().orc_virial_from_fshift(b, _p(fs), _p(vir))
    return vir.reshape(3, 3)


class _FepParams(C.Structure):
    _fields_ = [("rc", C.c_float), ("epsfac", C.c_float), ("k_rf", C.c_float), ("c_rf", C.c_float), ("disp_cpot", C.c_float),
                ("rep_cpot", C.c_float), ("lambda_coul", C.c_float), ("lambda_vdw", C.c_float), ("alpha_coul", C.c_float),
                ("alpha_vdw", C.c_float), ("lam_power", C.c_int), ("sigma6_def", C.c_float), ("sigma6_min", C.c_float),
                ("beta", C.c_float), ("sh_ewald", C.c_float), ("rvdw_switch", C.c_float),
                ("ljpme", C.c_int), ("ewaldcoeff_lj", C.c_float), ("sh_lj_ewald", C.c_float), ("rvdw", C.c_float)]


def fep_kernel(x, shift_vec, nbfp, typeA, typeB, qA, qB, iinr, shift, jindex, jjnr, excl_fep, rc, lambda_coul, lambda_vdw,
               epsfac=138.935458, k_rf=0.0, c_rf=0.0, sc_alpha=0.5, sc_power=1, sc_sigma=0.3, sc_sigma_min=0.3, sc_coul=False,
               ewaldcoeff=0.0, sh_ewald=0.0, rvdw_switch=0.0, ljpme=0, ewaldcoeff_lj=0.0, sh_lj_ewald=0.0, rvdw=0.0):
    """Plain-C restatement of the reference's free-energy kernel (orc_fep_kernel) on a perturbed pair list in t_nblist form; the
    soft-core parameters are the inputrec's (sc_alpha, sc_power, sc_sigma, sc_sigma_min, sc_coul), processed like
    interaction_const_t::SoftCoreParameters (mdtypes/interaction_const.cpp).  Returns f, fshift, (Vc, Vv, dvdl_coul, dvdl_vdw)."""
    x = _f32(x).reshape(-1, 3)
    n = x.shape[0]
    sv = _f32(shift_vec).reshape(45, 3)
    nb = _f32(nbfp).ravel()
    ntypes = int(round((nb.size // 2) ** 0.5))
    tA, tB, cA, cB = _i32(typeA), _i32(typeB), _f32(qA), _f32(qB)
    ii, sh, ji, jj = _i32(iinr), _i32(shift), _i32(jindex), _i32(jjnr)
    ex = np.ascontiguousarray(excl_fep, dtype=np.int8)
    # rvdw_switch > 0: LJ potential switch (no potential shift then, as interaction_const_t stores it)
    cpot6, cpot12 = (0.0, 0.0) if rvdw_switch > 0 else (-1.0 / (rvdw or rc) ** 6, -1.0 / (rvdw or rc) ** 12)
    p = _FepParams(rc, epsfac, k_rf, c_rf, cpot6, cpot12, lambda_coul, lambda_vdw,
                   sc_alpha if sc_coul else 0.0, sc_alpha, sc_power, sc_sigma ** 6, (sc_sigma_min ** 6) if sc_coul else 0.0,
                   ewaldcoeff, sh_ewald, rvdw_switch, int(ljpme), ewaldcoeff_lj, sh_lj_ewald, rvdw)
    f = np.zeros((n, 3), np.float32)
    fs = np.zeros((45, 3), np.float32)
    out = np.zeros(4, np.float32)
    lib().orc_fep_kernel(C.c_int(n), _p(x), _p(sv), C.c_int(ntypes), _p(nb), _p(tA), _p(tB), _p(cA), _p(cB), C.c_int(len(ii)), _p(ii),
                         _p(sh), _p(ji), _p(jj), _p(ex), C.byref(p), _p(f), _p(fs), _p(out))
    return f, fs, tuple(float(v) for v in out)


def fep_pair_list(x, box, rlist, perturbed, excl_off, excl_idx):
    """The perturbed pair list the reference's make_fep_list (nbnxm/pairlist.cpp:1699-1872) would hand to the free-energy kernel:
    every atom pair within rlist with at least one perturbed atom -- excluded pairs included, flagged 0 -- and every perturbed atom
    with itself (flag 0: its reaction-field self term), grouped by (i-atom, shift) in t_nblist form.  Returns iinr, shift, jindex,
    jjnr, excl_fep."""
    x = _f32(x)
    n = x.shape[0]
    pert = np.asarray(perturbed, bool)
    pairs = pair_set(x, box, rlist)  # no exclusions given: every pair in range
    pairs = pairs[pert[pairs[:, 0]] | pert[pairs[:, 1]]]
    flag = np.ones(len(pairs), np.int8)
    eo, ei = np.asarray(excl_off), np.asarray(excl_idx)
    exset = set()
    for a in np.nonzero(pert)[0]:
        for b in ei[eo[a]:eo[a + 1]]:
            exset.add((int(a), int(b)))
            exset.add((int(b), int(a)))
    for k, (i, j, sft) in enumerate(pairs):
        if sft == CENTRAL and (int(i), int(j)) in exset:
            flag[k] = 0
    selfp = np.nonzero(pert)[0].astype(np.int32)
    pairs = np.concatenate([pairs, np.stack([selfp, selfp, np.full(len(selfp), CENTRAL, np.int32)], 1)])
    flag = np.concatenate([flag, np.zeros(len(selfp), np.int8)])
    order = np.lexsort((pairs[:, 1], pairs[:, 2], pairs[:, 0]))
    pairs, flag = pairs[order], flag[order]
    key = pairs[:, 0].astype(np.int64) * 64 + pairs[:, 2]
    uk, start = np.unique(key, return_index=True)
    return ((uk // 64).astype(np.int32), (uk % 64).astype(np.int32), np.append(start, len(pairs)).astype(np.int32),
            np.ascontiguousarray(pairs[:, 1], dtype=np.int32), flag)


BONDED_KINDS = ("bonds", "angles", "urey_bradley", "pdihs", "rbdihs", "idihs", "pidihs", "lj14")
BONDED_NRAL = (2, 3, 3, 4, 4, 4, 4, 2)


def bonded(kind, iatoms, params6, x, q, box_matrix, epsfac_fudge=138.935458 * 0.5):
    """Plain-C restatement (orc_bonded) of the reference's GPU bonded kernels for one interaction type (`kind`: index into
    BONDED_KINDS).  iatoms[nbonds, nral + 1] = {parameter index, atoms}; params6[ntypes, 6]; box_matrix 3x3 lower-triangular.
    Returns f[n,3], fshift[45,3] (float64), (energy, Coulomb-14 energy for lj14)."""
    k = BONDED_KINDS.index(kind) if isinstance(kind, str) else int(kind)
    ia = _i32(iatoms).reshape(-1, BONDED_NRAL[k] + 1)
    p6 = _f32(params6).reshape(-1, 6)
    x = _f32(x).reshape(-1, 3)
    n = x.shape[0]
    q = _f32(q)
    b = _f32(box_matrix).reshape(9)
    f = np.zeros((n, 3), np.float64)
    fs = np.zeros((45, 3), np.float64)
    e = np.zeros(2, np.float64)
    L = lib()
    L.orc_bonded.restype = C.c_int
    rc = L.orc_bonded(C.c_int(k), C.c_int(len(ia)), _p(ia), _p(p6), C.c_int(n), _p(x), _p(q), _p(b), C.c_float(epsfac_fudge), _p(f), _p(fs),
                      _p(e))
    if rc != 0:
        raise ValueError("orc_bonded: bad input (%d)" % rc)
    return f, fs, (float(e[0]), float(e[1]))


def fep_list_canonical(lst, keep=None):
    """Sorted keys of the pairs of a perturbed pair list in t_nblist form, independent of which atom is listed as i and of how the
    entries are split: (lower atom, higher atom, shift seen from the lower atom, exclusion flag)."""
    ii, sh, ji, jj, ex = [np.asarray(a) for a in lst]
    n = np.diff(ji)
    i, s, j, e = np.repeat(ii, n).astype(np.int64), np.repeat(sh, n).astype(np.int64), jj.astype(np.int64), ex.astype(np.int64)
    if keep is not None:
        i, s, j, e = i[keep], s[keep], j[keep], e[keep]
    swap = i > j
    a, b, s = np.where(swap, j, i), np.where(swap, i, j), np.where(swap, SHIFTS - 1 - s, s)
    return np.sort((((a << 24) | b) << 8 | s) * 2 + e)


def fep_list_within(lst, x, box, rlist):
    """Mask of the pairs of a t_nblist within rlist (float32, x_i + shift - x_j as the reference's make_fep_list evaluates it,
    nbnxm/pairlist.cpp:1811-1824): the reference keeps pairs up to rlist + a cluster-size dependent buffer in its lists."""
    ii, sh, ji, jj, ex = [np.asarray(a) for a in lst]
    n = np.diff(ji)
    i, sft = np.repeat(ii, n), np.repeat(sh, n)
    x = _f32(x)
    d = x[jj] - (x[i] + shift_vectors(box)[sft]).astype(np.float32)
    r2 = ((d[:, 0] * d[:, 0]).astype(np.float32) + (d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32) + (d[:, 2] * d[:, 2]).astype(np.float32)
    return r2.astype(np.float32) < np.float32(np.float32(rlist) * np.float32(rlist))
