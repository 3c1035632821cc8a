"""ctypes binding of the C ABI declared in include/b200nb.h (libb200nb.so).

`NbnxmGpu` is the host-side handle corresponding to the reference's `NbnxmGpu*` / `nonbonded_verlet_t::gpu_nbv`
(src/gromacs/nbnxm/nbnxm.h:406, cuda/nbnxm_cuda_types.h:150-225); its methods are 1:1 with the C entry points.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIFTS, CENTRAL = 45, 22
EEL_CUT, EEL_RF, EEL_EWALD = 0, 1, 2
FLAG_ENERGY, FLAG_VIRIAL = 1, 2

EXPORTS = [
    "b200nb_create", "b200nb_destroy", "b200nb_last_error", "b200nb_stream", "b200nb_set_stream", "b200nb_synchronize",
    "b200nb_set_params", "b200nb_set_vdw", "b200nb_set_atoms", "b200nb_set_box", "b200nb_put_on_grid", "b200nb_build_pairlist",
    "b200nb_set_x", "b200nb_clear_outputs", "b200nb_launch_force", "b200nb_launch_prune", "b200nb_get_f",
    "b200nb_get_outputs", "b200nb_compute", "b200nb_step", "b200nb_dd_create_window", "b200nb_dd_open_peer", "b200nb_dd_set_plan", "b200nb_dd_set_links",
    "b200nb_dd_step", "b200nb_dd_status", "b200nb_halo_pack_x", "b200nb_halo_unpack_f", "b200nb_get_stats",
    "b200nb_get_grid_order", "b200nb_get_tiles", "b200nb_get_pairs", "b200nb_time_force_kernel", "b200nb_time_step",
    "b200nb_set_grid_atoms", "b200nb_upload_pairlist", "b200nb_copy_xq_grid", "b200nb_get_f_grid", "b200nb_set_shift_vec", "b200nb_set_ewald_table", "b200nb_describe",
    "b200nb_dd_wrap_classify", "b200nb_dd_select_lower_face", "b200nb_dd_partition_indices", "b200nb_dd_pack_atoms", "b200nb_dd_merge_home",
    "b200nb_dd_gather_int", "b200nb_dd_set_global_topology", "b200nb_dd_set_local_atoms", "b200nb_dd_wrap_classify_nd",
    "b200nb_dd_select_boundary", "b200nb_set_box_triclinic",
    "b200nb_fep_set_atoms", "b200nb_fep_upload_list", "b200nb_fep_launch", "b200nb_fep_get_outputs",
    "b200nb_fep_build_list", "b200nb_fep_get_list",
    "b200nb_bonded_set_list", "b200nb_bonded_launch", "b200nb_bonded_get_energies", "b200nb_bonded_in_step", "b200nb_fep_in_step", "b200nb_bonded_set_pbc",
]


class B200NBError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("ntypes", C.c_int), ("nbfp_host", C.c_void_p), ("rc", C.c_float), ("rlist_outer", C.c_float),
                ("rlist_inner", C.c_float), ("eeltype", C.c_int), ("epsfac", C.c_float), ("k_rf", C.c_float),
                ("c_rf", C.c_float), ("ewald_beta", C.c_float), ("sh_ewald", C.c_float), ("disp_cpot", C.c_float),
                ("rep_cpot", C.c_float), ("comb_rule", C.c_int), ("max_tiles_per_entry", C.c_int)]


class _Vdw(C.Structure):  # b200nb_vdw_t
    _fields_ = [("vdw_modifier", C.c_int), ("rvdw", C.c_float), ("rvdw_switch", C.c_float), ("disp_c2", C.c_float),
                ("disp_c3", C.c_float), ("rep_c2", C.c_float), ("rep_c3", C.c_float), ("sw_c3", C.c_float),
                ("sw_c4", C.c_float), ("sw_c5", C.c_float), ("ljpme_comb_rule", C.c_int), ("ewaldcoeff_lj", C.c_float),
                ("sh_lj_ewald", C.c_float)]


class _DdLink(C.Structure):  # b200nb_dd_link_t
    _fields_ = [("send_peer", C.c_int), ("nsend", C.c_int), ("send_idx_host", C.c_void_p), ("shift", C.c_float * 3),
                ("peer_halo_offset", C.c_int), ("recv_peer", C.c_int), ("nrecv", C.c_int), ("peer_entry_offset", C.c_int),
                ("fshift_index", C.c_int)]


DD_MAX_LINKS, DD_MAX_PEERS = 16, 16  # NB_DD_MAX_LINKS / NB_DD_MAX_PEERS (b200nb_internal.h)
VDW_POTSHIFT, VDW_FORCESWITCH, VDW_POTSWITCH = 0, 1, 2
LJPME_NONE, LJPME_GEOM, LJPME_LB = 0, 1, 2


def lj_ewald_shift(ewaldcoeff_lj, rvdw):
    """interaction_const_t::sh_lj_ewald for a potential-shift modifier (mdlib/forcerec.cpp:709-713)."""
    import math
    crc2 = (float(ewaldcoeff_lj) * float(rvdw)) ** 2
    return (math.exp(-crc2) * (1 + crc2 + 0.5 * crc2 * crc2) - 1) / float(rvdw) ** 6


def vdw_modifier_constants(modifier, rvdw, rvdw_switch):
    """The LJ modifier constants of interaction_const_t as init_interaction_const derives them
    (mdlib/forcerec.cpp:850-874; force_switch_constants :787-801, potential_switch_constants :803-816), rounded to
    float where the reference stores a `real`."""
    f32 = lambda v: float(np.float32(v))
    d = dict(disp_cpot=0.0, rep_cpot=0.0, disp_c2=0.0, disp_c3=0.0, rep_c2=0.0, rep_c3=0.0, sw_c3=0.0, sw_c4=0.0, sw_c5=0.0)
    rc, rsw = float(rvdw), float(rvdw_switch)
    if modifier == VDW_POTSHIFT:
        d["disp_cpot"], d["rep_cpot"] = -1.0 / rc ** 6, -1.0 / rc ** 12
    elif modifier == VDW_FORCESWITCH:
        for name, pw in (("disp", 6.0), ("rep", 12.0)):
            c2 = f32(((pw + 1) * rsw - (pw + 4) * rc) / (rc ** (pw + 2) * (rc - rsw) ** 2))
            c3 = f32(-((pw + 1) * rsw - (pw + 3) * rc) / (rc ** (pw + 2) * (rc - rsw) ** 3))
            d[name + "_c2"], d[name + "_c3"] = c2, c3
            d[name + "_cpot"] = f32(-rc ** (-pw) + pw * c2 / 3 * (rc - rsw) ** 3 + pw * c3 / 4 * (rc - rsw) ** 4)
    elif modifier == VDW_POTSWITCH:
        d["sw_c3"], d["sw_c4"], d["sw_c5"] = f32(-10 / (rc - rsw) ** 3), f32(15 / (rc - rsw) ** 4), f32(-6 / (rc - rsw) ** 5)
    else:
        raise B200NBError("unknown vdw modifier %r" % (modifier,))
    return d


class _Stats(C.Structure):
    _fields_ = [("natoms", C.c_int), ("natoms_padded", C.c_int), ("nclusters", C.c_int), ("ncx", C.c_int),
                ("ncy", C.c_int), ("ntiles_outer", C.c_longlong), ("ntiles_inner", C.c_longlong),
                ("nentries", C.c_longlong), ("comb_geometric", C.c_int), ("nlaunches", C.c_longlong), ("ntiles_packed", C.c_longlong),
                ("nentries_nonlocal", C.c_longlong)]


_lib = None


def library_path():
    # B200NB_LIBRARY: alternative build of the same C ABI (kernel tuning experiments under profiles/tools)
    return os.environ.get("B200NB_LIBRARY") or os.path.join(_HERE, "libb200nb.so")


class _FepParams(C.Structure):
    _fields_ = [("lambda_coul", C.c_float), ("lambda_vdw", C.c_float), ("sc_alpha", C.c_float), ("sc_power", C.c_int),
                ("sc_sigma", C.c_float), ("sc_sigma_min", C.c_float), ("sc_coul", C.c_int)]


def load_library():
    """Loads libb200nb.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise B200NBError("libb200nb.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or make -C gmxapi_b200/csrc); there is no CPU fallback")
    L = C.CDLL(path)
    vp, ci, cf, cll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
    L.b200nb_create.argtypes = [C.POINTER(vp), ci]
    L.b200nb_destroy.argtypes = [vp]
    L.b200nb_destroy.restype = None
    L.b200nb_last_error.argtypes = [vp]
    L.b200nb_last_error.restype = C.c_char_p
    L.b200nb_stream.argtypes = [vp]
    L.b200nb_stream.restype = vp
    L.b200nb_synchronize.argtypes = [vp]
    L.b200nb_set_stream.argtypes = [vp, vp]
    L.b200nb_set_params.argtypes = [vp, C.POINTER(_Params)]
    L.b200nb_set_vdw.argtypes = [vp, C.POINTER(_Vdw)]
    L.b200nb_set_atoms.argtypes = [vp, ci, vp, vp, vp, vp]
    L.b200nb_set_box.argtypes = [vp, vp, vp]
    L.b200nb_set_box_triclinic.argtypes = [vp, vp]
    L.b200nb_fep_set_atoms.argtypes = [vp, vp, vp, vp, vp]
    L.b200nb_fep_upload_list.argtypes = [vp, ci, vp, vp, vp, vp, vp]
    L.b200nb_fep_launch.argtypes = [vp, C.POINTER(_FepParams)]
    L.b200nb_fep_build_list.argtypes = [vp, vp, vp]
    L.b200nb_bonded_set_list.argtypes = [vp, ci, ci, vp, ci, vp]
    L.b200nb_bonded_launch.argtypes = [vp, ci, C.c_float]
    L.b200nb_bonded_get_energies.argtypes = [vp, vp]
    L.b200nb_bonded_in_step.argtypes = [vp, ci, C.c_float]
    L.b200nb_bonded_set_pbc.argtypes = [vp, vp, ci]
    L.b200nb_fep_in_step.argtypes = [vp, C.POINTER(_FepParams)]
    L.b200nb_fep_get_list.argtypes = [vp, vp, vp, vp, vp, vp]
    L.b200nb_fep_get_outputs.argtypes = [vp, vp]
    L.b200nb_put_on_grid.argtypes = [vp, ci, vp, vp, ci, ci, cf, vp, ci]
    L.b200nb_build_pairlist.argtypes = [vp]
    L.b200nb_set_x.argtypes = [vp, vp, ci, ci, ci]
    L.b200nb_clear_outputs.argtypes = [vp]
    L.b200nb_launch_force.argtypes = [vp, ci, ci]
    L.b200nb_launch_prune.argtypes = [vp, ci, ci, ci]
    L.b200nb_get_f.argtypes = [vp, vp, ci, ci, ci, ci]
    L.b200nb_get_outputs.argtypes = [vp, vp, vp]
    L.b200nb_compute.argtypes = [vp, vp, ci, vp, vp, vp]
    L.b200nb_step.argtypes = [vp, vp, ci, vp]
    L.b200nb_dd_create_window.argtypes = [vp, ci, ci, vp, C.POINTER(vp)]
    L.b200nb_dd_open_peer.argtypes = [vp, ci, vp, vp, ci]
    L.b200nb_dd_set_plan.argtypes = [vp, ci, ci, vp, ci, vp, ci]
    L.b200nb_dd_set_links.argtypes = [vp, ci, ci, ci, C.POINTER(_DdLink)]
    L.b200nb_dd_step.argtypes = [vp, vp, vp, ci]
    L.b200nb_dd_status.argtypes = [vp]
    L.b200nb_halo_pack_x.argtypes = [vp, vp, vp, ci, vp, vp]
    L.b200nb_halo_unpack_f.argtypes = [vp, vp, vp, ci, vp]
    L.b200nb_get_stats.argtypes = [vp, C.POINTER(_Stats)]
    L.b200nb_get_grid_order.argtypes = [vp, vp, ci]
    L.b200nb_get_tiles.argtypes = [vp, ci, vp, cll]
    L.b200nb_get_tiles.restype = cll
    L.b200nb_get_pairs.argtypes = [vp, cf, vp, cll]
    L.b200nb_get_pairs.restype = cll
    L.b200nb_time_force_kernel.argtypes = [vp, ci, ci, ci, ci, ci, C.POINTER(cf)]
    L.b200nb_time_step.argtypes = [vp, vp, vp, ci, ci, ci, ci, C.POINTER(cf), C.POINTER(cf)]
    L.b200nb_set_grid_atoms.argtypes = [vp, ci, vp, vp]
    L.b200nb_upload_pairlist.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci]
    L.b200nb_copy_xq_grid.argtypes = [vp, vp, ci, ci]
    L.b200nb_get_f_grid.argtypes = [vp, vp, ci, ci]
    L.b200nb_set_shift_vec.argtypes = [vp, vp]
    L.b200nb_set_ewald_table.argtypes = [vp, vp, ci, cf]
    L.b200nb_describe.argtypes = [vp, C.c_char_p, ci]
    L.b200nb_dd_wrap_classify.argtypes = [vp, vp, ci, vp, vp, ci, ci, vp]
    L.b200nb_dd_select_lower_face.argtypes = [vp, vp, ci, cf, cf, vp]
    L.b200nb_dd_partition_indices.argtypes = [vp, vp, ci, ci, vp, vp]
    L.b200nb_dd_pack_atoms.argtypes = [vp, vp, ci, vp, vp, vp]
    L.b200nb_dd_merge_home.argtypes = [vp, vp, ci, vp, vp, vp, ci, vp, vp]
    L.b200nb_dd_gather_int.argtypes = [vp, vp, ci, vp, vp]
    L.b200nb_dd_set_global_topology.argtypes = [vp, ci, vp, vp, vp, vp]
    L.b200nb_dd_set_local_atoms.argtypes = [vp, vp, ci]
    L.b200nb_dd_wrap_classify_nd.argtypes = [vp, vp, ci, vp, vp, vp, vp]
    L.b200nb_dd_select_boundary.argtypes = [vp, vp, ci, vp, vp, vp, cf, vp]
    _lib = L
    return L


BONDED_KINDS = ("bonds", "angles", "urey_bradley", "pdihs", "rbdihs", "idihs", "pidihs", "lj14")  # B200NB_BONDED_*
BONDED_NRAL = (2, 3, 3, 4, 4, 4, 4, 2)


def _ptr(a):
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class NbnxmGpu:
    """One nonbonded context on one GPU (the reference's NbnxmGpu, created by Nbnxm::gpu_init)."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.b200nb_create(C.byref(h), int(device))
        if rc != 0:
            raise B200NBError("b200nb_create failed (rc=%d): no usable CUDA device %d -- this path has no CPU "
                              "fallback" % (rc, device))
        self._h = h
        self.natoms = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200nb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._L.b200nb_last_error(self._h)
            raise B200NBError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))

    @property
    def stream(self):
        return self._L.b200nb_stream(self._h)

    def set_stream(self, cuda_stream):
        """Issue all work on the caller's CUDA stream (an integer handle, e.g. torch.cuda.Stream().cuda_stream)."""
        self._check(self._L.b200nb_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else 0)), "set_stream")

    def synchronize(self):
        self._check(self._L.b200nb_synchronize(self._h), "synchronize")

    def set_params(self, nbfp, rc, rlist_outer=None, rlist_inner=None, eeltype=EEL_CUT, epsfac=138.935458, k_rf=0.0,
                   c_rf=0.0, ewald_beta=0.0, sh_ewald=0.0, disp_cpot=None, rep_cpot=None, comb_rule=0,
                   max_tiles_per_entry=0):
        nbfp = np.ascontiguousarray(nbfp, dtype=np.float32).ravel()
        ntypes = int(round((nbfp.size // 2) ** 0.5))
        if ntypes * ntypes * 2 != nbfp.size:
            raise B200NBError("nbfp must hold ntypes*ntypes*2 values")
        rlo = float(rlist_outer) if rlist_outer else float(rc)
        rli = float(rlist_inner) if rlist_inner else rlo
        p = _Params(ntypes, nbfp.ctypes.data, rc, rlo, rli, eeltype, epsfac, k_rf, c_rf, ewald_beta, sh_ewald,
                    -1.0 / rc ** 6 if disp_cpot is None else disp_cpot,
                    -1.0 / rc ** 12 if rep_cpot is None else rep_cpot, comb_rule, max_tiles_per_entry)
        self._check(self._L.b200nb_set_params(self._h, C.byref(p)), "set_params")

    def set_ewald_table(self, table_f, scale):
        """Tabulated Ewald force correction (EwaldCorrectionTables::tableF, tableScale); table_f=None returns to the analytical one."""
        if table_f is None:
            self._check(self._L.b200nb_set_ewald_table(self._h, None, 0, 0.0), "set_ewald_table")
            return
        t = np.ascontiguousarray(table_f, dtype=np.float32)
        self._check(self._L.b200nb_set_ewald_table(self._h, _ptr(t), len(t), float(scale)), "set_ewald_table")

    def set_vdw(self, vdw_modifier=VDW_POTSHIFT, rvdw=0.0, rvdw_switch=0.0, constants=None, ljpme=LJPME_NONE,
                ewaldcoeff_lj=0.0, sh_lj_ewald=0.0):
        """b200nb_set_vdw: LJ force / potential switch, VdW cut-off (<= rc) and the LJ-PME grid correction.  `constants`: the
        c2/c3/c3..c5 values (vdw_modifier_constants); the matching potential shifts belong in set_params(disp_cpot=, rep_cpot=)."""
        k = constants or {}
        v = _Vdw(int(vdw_modifier), float(rvdw), float(rvdw_switch), k.get("disp_c2", 0.0), k.get("disp_c3", 0.0),
                 k.get("rep_c2", 0.0), k.get("rep_c3", 0.0), k.get("sw_c3", 0.0), k.get("sw_c4", 0.0), k.get("sw_c5", 0.0),
                 int(ljpme), float(ewaldcoeff_lj), float(sh_lj_ewald))
        self._check(self._L.b200nb_set_vdw(self._h, C.byref(v)), "set_vdw")

    def set_atoms(self, types, q, excl_off=None, excl_idx=None):
        types = np.ascontiguousarray(types, dtype=np.int32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        eo = np.ascontiguousarray(excl_off, dtype=np.int32) if excl_off is not None else None
        ei = np.ascontiguousarray(excl_idx, dtype=np.int32) if excl_idx is not None else None
        self.natoms = int(types.shape[0])
        self._check(self._L.b200nb_set_atoms(self._h, self.natoms, _ptr(types), _ptr(q), _ptr(eo), _ptr(ei)),
                    "set_atoms")

    def set_box(self, box, pbc=(1, 1, 1)):
        """box: 3 edge lengths (rectangular) or the 3 x 3 lower-triangular box matrix of a triclinic cell (fully periodic)"""
        b = np.ascontiguousarray(box, dtype=np.float32)
        if b.size == 9:
            m = b.reshape(3, 3)
            if m[1, 0] != 0 or m[2, 0] != 0 or m[2, 1] != 0:
                self._check(self._L.b200nb_set_box_triclinic(self._h, _ptr(np.ascontiguousarray(m.ravel()))), "set_box_triclinic")
                return
            b = np.ascontiguousarray(np.diag(m))
        p = np.ascontiguousarray(pbc, dtype=np.int32)
        self._check(self._L.b200nb_set_box(self._h, _ptr(b), _ptr(p)), "set_box")

    def put_on_grid(self, x, lower, upper, grid_index=0, atom_begin=0, atom_end=None, density=0.0, on_device=False):
        lo = np.ascontiguousarray(lower, dtype=np.float32)
        up = np.ascontiguousarray(upper, dtype=np.float32)
        if atom_end is None:
            atom_end = self.natoms
        if not on_device:
            x = np.ascontiguousarray(x, dtype=np.float32)
        self._check(self._L.b200nb_put_on_grid(self._h, grid_index, _ptr(lo), _ptr(up), atom_begin, atom_end,
                                               float(density), _ptr(x), int(on_device)), "put_on_grid")

    def build_pairlist(self):
        self._check(self._L.b200nb_build_pairlist(self._h), "build_pairlist")

    def set_x(self, x, on_device=False, atom_begin=0, atom_end=None):
        if atom_end is None:
            atom_end = self.natoms
        if not on_device:
            x = np.ascontiguousarray(x, dtype=np.float32)
        self._check(self._L.b200nb_set_x(self._h, _ptr(x), int(on_device), atom_begin, atom_end), "set_x")

    def clear_outputs(self):
        self._check(self._L.b200nb_clear_outputs(self._h), "clear_outputs")

    def launch_force(self, locality=-1, flags=0):
        self._check(self._L.b200nb_launch_force(self._h, locality, flags), "launch_force")

    def launch_prune(self, locality=-1, part=0, num_parts=1):
        self._check(self._L.b200nb_launch_prune(self._h, locality, part, num_parts), "launch_prune")

    def get_f(self, f=None, on_device=False, accumulate=False, atom_begin=0, atom_end=None):
        if atom_end is None:
            atom_end = self.natoms
        if not on_device:
            if f is None:
                f = np.zeros((self.natoms, 3), np.float32)
            assert f.dtype == np.float32 and f.flags.c_contiguous
        self._check(self._L.b200nb_get_f(self._h, _ptr(f), int(on_device), int(accumulate), atom_begin, atom_end),
                    "get_f")
        return f

    def get_outputs(self):
        fs = np.zeros((SHIFTS, 3), np.float32)
        e = np.zeros(2, np.float64)
        self._check(self._L.b200nb_get_outputs(self._h, _ptr(fs), _ptr(e)), "get_outputs")
        return fs, float(e[0]), float(e[1])

    def compute(self, x, flags=0, f=None):
        """GmxForceCalculator::compute: returns (f, fshift, e_lj, e_el).  x / f that live in pinned host memory
        (e.g. torch pin_memory, cudaHostRegister) are read / written in place by the kernels."""
        if not (isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags.c_contiguous):
            x = np.ascontiguousarray(x, dtype=np.float32)
        if f is None:
            f = np.empty((self.natoms, 3), np.float32)
        elif not (f.dtype == np.float32 and f.flags.c_contiguous and f.size == 3 * self.natoms):
            raise B200NBError("compute: forces must be a C-contiguous float32 array of natoms*3")
        if x.size != 3 * self.natoms:
            raise B200NBError("compute: coordinates must hold natoms*3 values")
        if flags:
            fs = np.zeros((SHIFTS, 3), np.float32)
            e = np.zeros(2, np.float64)
            rc = self._L.b200nb_compute(self._h, x.ctypes.data, flags, f.ctypes.data, fs.ctypes.data, e.ctypes.data)
        else:
            fs, e = None, (0.0, 0.0)
            rc = self._L.b200nb_compute(self._h, x.ctypes.data, 0, f.ctypes.data, None, None)
        if rc != 0:
            self._check(rc, "compute")
        return f, fs, float(e[0]), float(e[1])

    def step(self, x_dev, f_dev, flags=0):
        """Device-resident step (x_dev, f_dev: device addresses of natoms*3 floats); asynchronous."""
        self._check(self._L.b200nb_step(self._h, _ptr(x_dev), flags, _ptr(f_dev)), "step")

    # ---- domain-decomposed step over peer-memory halo windows ----
    def dd_create_window(self, max_halo, max_send):
        """Returns (64-byte CUDA IPC handle, device pointer) of this rank's halo window."""
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        self._check(self._L.b200nb_dd_create_window(self._h, int(max_halo), int(max_send), handle, C.byref(ptr)),
                    "dd_create_window")
        return handle.raw, int(ptr.value)

    def dd_open_peer(self, side, ipc_handle=None, window_ptr=None, peer_max_halo=0):
        hb = C.create_string_buffer(ipc_handle, 64) if ipc_handle is not None else None
        self._check(self._L.b200nb_dd_open_peer(self._h, int(side), hb, C.c_void_p(window_ptr) if window_ptr else None,
                                                int(peer_max_halo)), "dd_open_peer")

    def dd_set_plan(self, nhome, nhalo, send_idx, shift, halo_fshift_index=-1):
        """1-D plan: one link (send to peer 0, receive from peer 1); halo_fshift_index: shift-force slot that also gets the
        forces computed here on the halo atoms when those arrived across the periodic edge, else -1."""
        si = np.ascontiguousarray(send_idx, dtype=np.int32)
        sh = np.ascontiguousarray(shift, dtype=np.float32)
        self._check(self._L.b200nb_dd_set_plan(self._h, int(nhome), int(nhalo), _ptr(si) if si.size else None, int(si.size),
                                               _ptr(sh), int(halo_fshift_index)), "dd_set_plan")

    # -- perturbed (free-energy) pairs: csrc/fep.cu -------------------------------------------------------------------------------
    def fep_set_atoms(self, typeA, typeB, qA, qB):
        """A / B state of every atom (t_mdatoms::typeA / typeB / chargeA / chargeB), atom order"""
        tA, tB = np.ascontiguousarray(typeA, dtype=np.int32), np.ascontiguousarray(typeB, dtype=np.int32)
        cA, cB = np.ascontiguousarray(qA, dtype=np.float32), np.ascontiguousarray(qB, dtype=np.float32)
        self._check(self._L.b200nb_fep_set_atoms(self._h, _ptr(tA), _ptr(tB), _ptr(cA), _ptr(cB)), "fep_set_atoms")

    def fep_upload_list(self, iinr, shift, jindex, jjnr, excl_fep):
        """the perturbed pair list in t_nblist form (mdtypes/nblist.h): i-entries {iinr, shift, jindex}, jjnr, excl_fep"""
        ii, sh = np.ascontiguousarray(iinr, dtype=np.int32), np.ascontiguousarray(shift, dtype=np.int32)
        ji, jj = np.ascontiguousarray(jindex, dtype=np.int32), np.ascontiguousarray(jjnr, dtype=np.int32)
        ex = np.ascontiguousarray(excl_fep, dtype=np.int8)
        self._check(self._L.b200nb_fep_upload_list(self._h, int(len(ii)), _ptr(ii), _ptr(sh), _ptr(ji), _ptr(jj), _ptr(ex)), "fep_upload_list")

    def fep_build_list(self):
        """the perturbed pair list from the gridded coordinates, on the device (make_fep_list's output); returns (nri, nrj)"""
        nri, nrj = C.c_int(0), C.c_int(0)
        self._check(self._L.b200nb_fep_build_list(self._h, C.byref(nri), C.byref(nrj)), "fep_build_list")
        self._fep_sizes = (nri.value, nrj.value)
        return self._fep_sizes

    def fep_list(self):
        """the list fep_build_list made, on the host: iinr, shift, jindex, jjnr, excl_fep"""
        nri, nrj = self._fep_sizes
        ii, sh, ji = np.zeros(nri, np.int32), np.zeros(nri, np.int32), np.zeros(nri + 1, np.int32)
        jj, ex = np.zeros(nrj, np.int32), np.zeros(nrj, np.int8)
        self._check(self._L.b200nb_fep_get_list(self._h, _ptr(ii), _ptr(sh), _ptr(ji), _ptr(jj), _ptr(ex)), "fep_get_list")
        return ii, sh, ji, jj, ex

    # -- listed ("bonded") interactions: csrc/bonded.cu --------------------------------------------------------------------------
    def bonded_set_list(self, kind, iatoms, params6):
        """one interaction type of gmx::GpuBonded: iatoms[nbonds, nral + 1] = {parameter index, atoms}, params6[nparams, 6]"""
        k = BONDED_KINDS.index(kind) if isinstance(kind, str) else int(kind)
        ia = np.ascontiguousarray(iatoms, dtype=np.int32).reshape(-1, BONDED_NRAL[k] + 1)
        p6 = np.ascontiguousarray(params6, dtype=np.float32).reshape(-1, 6)
        self._check(self._L.b200nb_bonded_set_list(self._h, k, len(ia), _ptr(ia), len(p6), _ptr(p6)), "bonded_set_list")

    def bonded_launch(self, flags=0, epsfac_fudge=138.935458 * 0.5):
        self._check(self._L.b200nb_bonded_launch(self._h, int(flags), float(epsfac_fudge)), "bonded_launch")

    def bonded_set_pbc(self, box_matrix=None, npbcdim=3):
        """GpuBonded::setPbc: the cell of the bonded image search when it is not the context's (None: the context's)"""
        b = None if box_matrix is None else np.ascontiguousarray(box_matrix, dtype=np.float32).reshape(9)
        self._check(self._L.b200nb_bonded_set_pbc(self._h, _ptr(b), int(npbcdim)), "bonded_set_pbc")

    def bonded_in_step(self, enable=True, epsfac_fudge=138.935458 * 0.5):
        """make the bonded kernel part of step() / compute() (inside the captured step graph)"""
        self._check(self._L.b200nb_bonded_in_step(self._h, int(bool(enable)), float(epsfac_fudge)), "bonded_in_step")

    def bonded_energies(self):
        """{kind: energy} + "coul14", summed over the launches since the last call (read and reset)"""
        out = np.zeros(len(BONDED_KINDS) + 1, np.float64)
        self._check(self._L.b200nb_bonded_get_energies(self._h, _ptr(out)), "bonded_get_energies")
        return dict(zip(BONDED_KINDS + ("coul14",), (float(v) for v in out)))

    def fep_launch(self, lambda_coul, lambda_vdw, sc_alpha=0.5, sc_power=1, sc_sigma=0.3, sc_sigma_min=0.3, sc_coul=False):
        p = _FepParams(lambda_coul, lambda_vdw, sc_alpha, int(sc_power), sc_sigma, sc_sigma_min, int(bool(sc_coul)))
        self._check(self._L.b200nb_fep_launch(self._h, C.byref(p)), "fep_launch")

    def fep_in_step(self, lambda_coul=None, lambda_vdw=None, sc_alpha=0.5, sc_power=1, sc_sigma=0.3, sc_sigma_min=0.3, sc_coul=False):
        """make the free-energy kernel part of step() / compute() with these parameters (lambda_coul=None: take it out again)"""
        if lambda_coul is None:
            self._check(self._L.b200nb_fep_in_step(self._h, None), "fep_in_step")
            return
        p = _FepParams(lambda_coul, lambda_vdw, sc_alpha, int(sc_power), sc_sigma, sc_sigma_min, int(bool(sc_coul)))
        self._check(self._L.b200nb_fep_in_step(self._h, C.byref(p)), "fep_in_step")

    def fep_outputs(self):
        """(Vc, Vv, dV/dlambda_coul, dV/dlambda_vdw) of the launches since the last call"""
        out = np.zeros(4, np.float64)
        self._check(self._L.b200nb_fep_get_outputs(self._h, _ptr(out)), "fep_get_outputs")
        return tuple(float(v) for v in out)

    # -- repartitioning on the device (b200nb_dd_* of csrc/dd_partition.cu); pointers are device addresses (ints) ----------------
    def dd_wrap_classify(self, x_dev, n, box, bounds, nranks, rank, code_dev):
        b = np.ascontiguousarray(box, dtype=np.float32)
        bd = np.ascontiguousarray(bounds, dtype=np.float32)
        self._check(self._L.b200nb_dd_wrap_classify(self._h, _ptr(x_dev), int(n), _ptr(b), _ptr(bd), int(nranks), int(rank), _ptr(code_dev)),
                    "dd_wrap_classify")

    def dd_wrap_classify_nd(self, x_dev, n, box, grid, coords, code_dev):
        b = np.ascontiguousarray(box, dtype=np.float32)
        g = np.ascontiguousarray(grid, dtype=np.int32)
        c = np.ascontiguousarray(coords, dtype=np.int32)
        self._check(self._L.b200nb_dd_wrap_classify_nd(self._h, _ptr(x_dev), int(n), _ptr(b), _ptr(g), _ptr(c), _ptr(code_dev)),
                    "dd_wrap_classify_nd")

    def dd_select_boundary(self, x_dev, n, lo, hi, offset, rlist, code_dev):
        lo = np.ascontiguousarray(lo, dtype=np.float32)
        hi = np.ascontiguousarray(hi, dtype=np.float32)
        o = np.ascontiguousarray(offset, dtype=np.int32)
        self._check(self._L.b200nb_dd_select_boundary(self._h, _ptr(x_dev), int(n), _ptr(lo), _ptr(hi), _ptr(o), float(rlist), _ptr(code_dev)),
                    "dd_select_boundary")

    def dd_select_lower_face(self, x_dev, n, lo, rlist, code_dev):
        self._check(self._L.b200nb_dd_select_lower_face(self._h, _ptr(x_dev), int(n), float(lo), float(rlist), _ptr(code_dev)),
                    "dd_select_lower_face")

    def dd_partition_indices(self, code_dev, n, ncodes, idx_dev):
        """Stable partition of 0 .. n-1 by code; returns the counts per code (list of ncodes ints)."""
        cnt = np.zeros(32, np.int32)
        self._check(self._L.b200nb_dd_partition_indices(self._h, _ptr(code_dev), int(n), int(ncodes), _ptr(idx_dev), _ptr(cnt)),
                    "dd_partition_indices")
        return [int(v) for v in cnt[:ncodes]]

    def dd_pack_atoms(self, idx_dev, m, gid_dev, x_dev, out4_dev):
        self._check(self._L.b200nb_dd_pack_atoms(self._h, _ptr(idx_dev), int(m), _ptr(gid_dev), _ptr(x_dev), _ptr(out4_dev)), "dd_pack_atoms")

    def dd_merge_home(self, stay_idx_dev, nstay, gid_dev, x_dev, arrived4_dev, narrived, gid_out_dev, x_out_dev):
        self._check(self._L.b200nb_dd_merge_home(self._h, _ptr(stay_idx_dev), int(nstay), _ptr(gid_dev), _ptr(x_dev), _ptr(arrived4_dev),
                                                 int(narrived), _ptr(gid_out_dev), _ptr(x_out_dev)), "dd_merge_home")

    def dd_gather_int(self, idx_dev, m, in_dev, out_dev):
        self._check(self._L.b200nb_dd_gather_int(self._h, _ptr(idx_dev), int(m), _ptr(in_dev), _ptr(out_dev)), "dd_gather_int")

    def dd_set_global_topology(self, types, q, excl_off=None, excl_idx=None):
        types = np.ascontiguousarray(types, dtype=np.int32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        eo = np.ascontiguousarray(excl_off, dtype=np.int32) if excl_off is not None else None
        ei = np.ascontiguousarray(excl_idx, dtype=np.int32) if excl_idx is not None else None
        self._check(self._L.b200nb_dd_set_global_topology(self._h, int(types.shape[0]), _ptr(types), _ptr(q), _ptr(eo), _ptr(ei)),
                    "dd_set_global_topology")

    def dd_set_local_atoms(self, local_gid_dev, nlocal):
        self.natoms = int(nlocal)
        self._check(self._L.b200nb_dd_set_local_atoms(self._h, _ptr(local_gid_dev), int(nlocal)), "dd_set_local_atoms")

    def dd_set_links(self, nhome, nhalo, links):
        """General plan (b200nb_dd_set_links).  links: list of dicts with keys send_peer, send_idx (int32 array), shift[3],
        peer_halo_offset, recv_peer, nrecv, peer_entry_offset, fshift_index."""
        arr = (_DdLink * max(len(links), 1))()
        keep = []
        for k, l in enumerate(links):
            si = np.ascontiguousarray(l["send_idx"], dtype=np.int32)
            keep.append(si)
            arr[k].send_peer, arr[k].nsend = int(l["send_peer"]), int(si.size)
            arr[k].send_idx_host = si.ctypes.data if si.size else None
            for d in range(3):
                arr[k].shift[d] = float(l["shift"][d])
            arr[k].peer_halo_offset = int(l["peer_halo_offset"])
            arr[k].recv_peer, arr[k].nrecv = int(l["recv_peer"]), int(l["nrecv"])
            arr[k].peer_entry_offset, arr[k].fshift_index = int(l["peer_entry_offset"]), int(l["fshift_index"])
        self._check(self._L.b200nb_dd_set_links(self._h, int(nhome), int(nhalo), len(links), arr), "dd_set_links")

    def dd_step(self, x_home, f_home, flags=0):
        """x_home / f_home: addresses (device or pinned host) of nhome*3 floats; asynchronous."""
        rc = self._L.b200nb_dd_step(self._h, x_home, f_home, flags)
        if rc != 0:
            self._check(rc, "dd_step")

    def dd_status(self):
        self._check(self._L.b200nb_dd_status(self._h), "dd_status")

    def halo_pack_x(self, x_dev, index_dev, n, shift, out_dev):
        s = np.ascontiguousarray(shift, dtype=np.float32)
        self._check(self._L.b200nb_halo_pack_x(self._h, _ptr(x_dev), _ptr(index_dev), n, _ptr(s), _ptr(out_dev)),
                    "halo_pack_x")

    def halo_unpack_f(self, f_dev, index_dev, n, in_dev):
        self._check(self._L.b200nb_halo_unpack_f(self._h, _ptr(f_dev), _ptr(index_dev), n, _ptr(in_dev)),
                    "halo_unpack_f")

    def describe(self):
        """One line about this context (device, atoms, grid, kernel flavour, list sizes) for logs."""
        buf = C.create_string_buffer(1024)
        self._check(self._L.b200nb_describe(self._h, buf, 1024), "describe")
        return buf.value.decode()

    def stats(self):
        s = _Stats()
        self._check(self._L.b200nb_get_stats(self._h, C.byref(s)), "get_stats")
        return {k: getattr(s, k) for k, _ in _Stats._fields_}

    def grid_order(self):
        n = self.stats()["natoms_padded"]
        out = np.zeros(n, np.int32)
        rc = self._L.b200nb_get_grid_order(self._h, _ptr(out), n)
        if rc < 0:
            self._check(rc, "get_grid_order")
        return out

    def tiles(self, outer=False):
        n = self._L.b200nb_get_tiles(self._h, int(outer), _ptr(None), 0)
        if n < 0:
            self._check(int(n), "get_tiles")
        out = np.zeros((max(n, 1), 3), np.int32)
        self._L.b200nb_get_tiles(self._h, int(outer), _ptr(out), n)
        return out[:n]

    def pairs(self, r):
        n = self._L.b200nb_get_pairs(self._h, float(r), _ptr(None), 0)
        if n < 0:
            self._check(int(n), "get_pairs")
        out = np.zeros((max(n, 1), 3), np.int32)
        m = self._L.b200nb_get_pairs(self._h, float(r), _ptr(out), n)
        assert m == n
        return out[:n]

    def pair_count(self, r):
        n = self._L.b200nb_get_pairs(self._h, float(r), _ptr(None), 0)
        if n < 0:
            self._check(int(n), "get_pairs")
        return int(n)

    # -- reference-built grid and list (what the Nbnxm::gpu_* shim calls) -----------------------------------------------
    def set_grid_atoms(self, xq_grid, type_grid):
        """nbat->x() (nslots x 4) and nbat->params().type (nslots) in grid order (gpu_init_atomdata)."""
        xq = np.ascontiguousarray(xq_grid, dtype=np.float32).reshape(-1, 4)
        ty = np.ascontiguousarray(type_grid, dtype=np.int32)
        self._nslots = len(ty)
        self._check(self._L.b200nb_set_grid_atoms(self._h, len(ty), _ptr(xq), _ptr(ty)), "set_grid_atoms")

    def upload_pairlist(self, locality, sci, cj4, excl):
        """NbnxnPairlistGpu arrays as int32 / uint32 numpy arrays: sci (n, 4), cj4 (n, 8), excl (n, 32) (gpu_init_pairlist)."""
        sci = np.ascontiguousarray(sci, dtype=np.int32).reshape(-1, 4)
        cj4 = np.ascontiguousarray(cj4, dtype=np.int32).reshape(-1, 8)
        excl = np.ascontiguousarray(excl, dtype=np.uint32).reshape(-1, 32)
        self._check(self._L.b200nb_upload_pairlist(self._h, int(locality), _ptr(sci) if len(sci) else None, len(sci),
                                                   _ptr(cj4) if len(cj4) else None, len(cj4), _ptr(excl), len(excl)), "upload_pairlist")

    def copy_xq_grid(self, xq_grid, slot_begin=0, slot_end=None):
        xq = np.ascontiguousarray(xq_grid, dtype=np.float32).reshape(-1, 4)
        self._check(self._L.b200nb_copy_xq_grid(self._h, _ptr(xq), slot_begin, len(xq) if slot_end is None else slot_end), "copy_xq_grid")
        self.synchronize()  # xq may be a temporary

    def get_f_grid(self):
        f = np.zeros((self._nslots, 3), np.float32)
        self._check(self._L.b200nb_get_f_grid(self._h, _ptr(f), 0, self._nslots), "get_f_grid")
        self.synchronize()
        return f

    def time_step(self, x_dev, f_dev, flags=0, nwarm=3, niter=20, flush_l2=True):
        """(ms per device-resident step, ms of the force kernel inside it), CUDA-event timed on the context's stream."""
        a, b = C.c_float(), C.c_float()
        self._check(self._L.b200nb_time_step(self._h, _ptr(x_dev), _ptr(f_dev), flags, nwarm, niter, int(flush_l2),
                                             C.byref(a), C.byref(b)), "time_step")
        return a.value, b.value

    def time_force_kernel(self, locality=-1, flags=0, nwarm=3, niter=20, flush_l2=True):
        ms = C.c_float()
        self._check(self._L.b200nb_time_force_kernel(self._h, locality, flags, nwarm, niter, int(flush_l2),
                                                     C.byref(ms)), "time_force_kernel")
        return ms.value
