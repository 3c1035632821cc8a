"""Synthetic benchmark systems of the sizes BASELINE.json names.

The reference builds its benchmark boxes by tiling 1000 equilibrated SPC/E molecules
(src/gromacs/nbnxm/benchmark/bench_system.cpp:90-151, box 3.10736 nm per 3000 atoms).  We cannot ship
those coordinates, so we generate boxes of the SAME density, composition, force-field parameters
(bench_system.cpp:63-82) and exclusion topology (:192-195) from a seeded lattice with random molecular
orientations: molecule centres on a simple-cubic lattice of spacing 3.10736/10 nm, jittered.
"""
import numpy as np

SPACING = 3.10736 / 10.0  # 1000 molecules per 3.10736^3 nm^3
Q_O, Q_H = -0.8476, 0.4238
C6_O, C12_O = 0.0026173456, 2.634129e-06
ONE_4PI_EPS0 = 138.935458  # src/gromacs/math/units.h


class System:
    """Plain container: x[n,3] f32, box[3], types[n] i32, q[n] f32, nbfp[ntypes,ntypes,2] (6*C6, 12*C12),
    exclusions as CSR (excl_off[n+1], excl_idx) with self included, mol_id[n]."""

    def __init__(self, x, box, types, q, nbfp, excl_off, excl_idx, mol_id, name):
        self.x, self.box, self.types, self.q, self.nbfp = x, box, types, q, nbfp
        self.excl_off, self.excl_idx, self.mol_id, self.name = excl_off, excl_idx, mol_id, name
        self.n = x.shape[0]
        # triclinic cells: `box` is the diagonal of the lower-triangular box matrix, box_offdiag = box[YY][XX], box[ZZ][XX],
        # box[ZZ][YY] (all zero: rectangular)
        self.box_offdiag = np.zeros(3, np.float32)

    @property
    def box_matrix(self):
        b, o = self.box, self.box_offdiag
        return np.array([[b[0], 0, 0], [o[0], b[1], 0], [o[1], o[2], b[2]]], np.float32)


def water_box(nx, ny, nz, seed=20261017, jitter=0.02):
    """SPC/E-like water: nx*ny*nz molecules, 3 atoms each, box = SPACING*(nx,ny,nz)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nmol = nx * ny * nz
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3)
    centre = (g + 0.5) * SPACING + rng.uniform(-jitter, jitter, size=(nmol, 3))
    # random rotation per molecule from a random unit quaternion
    qt = rng.normal(size=(nmol, 4))
    qt /= np.linalg.norm(qt, axis=1, keepdims=True)
    w, a, b, c = qt.T
    R = np.stack([1 - 2 * (b * b + c * c), 2 * (a * b - c * w), 2 * (a * c + b * w),
                  2 * (a * b + c * w), 1 - 2 * (a * a + c * c), 2 * (b * c - a * w),
                  2 * (a * c - b * w), 2 * (b * c + a * w), 1 - 2 * (a * a + b * b)], -1).reshape(nmol, 3, 3)
    half = np.deg2rad(109.47) / 2
    local = np.array([[0.0, 0.0, 0.0],
                      [0.1 * np.sin(half), 0.0, 0.1 * np.cos(half)],
                      [-0.1 * np.sin(half), 0.0, 0.1 * np.cos(half)]])
    x = centre[:, None, :] + np.einsum("mij,aj->mai", R, local)
    box = np.array([nx, ny, nz], np.float64) * SPACING
    x = np.mod(x.reshape(-1, 3), box)  # put_atoms_in_box, bench_system.cpp:157
    x = x.astype(np.float32)
    box32 = box.astype(np.float32)
    x = np.where(x >= box32, x - box32, x).astype(np.float32)  # float rounding can land exactly on the edge
    n = 3 * nmol
    types = np.tile(np.array([0, 1, 1], np.int32), nmol)
    q = np.tile(np.array([Q_O, Q_H, Q_H], np.float32), nmol)
    nbfp = np.zeros((2, 2, 2), np.float32)
    nbfp[0, 0] = (6.0 * C6_O, 12.0 * C12_O)
    first = (np.arange(n) // 3) * 3
    excl_idx = (first[:, None] + np.arange(3)[None, :]).astype(np.int32).ravel()
    excl_off = (np.arange(n + 1) * 3).astype(np.int32)
    mol_id = (np.arange(n) // 3).astype(np.int32)
    return System(x, box32, types, q, nbfp, excl_off, excl_idx, mol_id, "water_%dx%dx%d" % (nx, ny, nz))


def put_atoms_in_triclinic_box(x, box_matrix):
    """put_atoms_in_box for a triclinic cell (pbcutil/pbc.cpp): from z down to x, whole box vectors are added / subtracted until
    0 <= x[d] < box[d][d] -- the atoms end up in the brick spanned by the diagonal, which is what the nbnxm grid covers."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3).copy()
    m = np.asarray(box_matrix, dtype=np.float32).reshape(3, 3)
    for d in (2, 1, 0):
        while True:
            lo = x[:, d] < 0
            hi = x[:, d] >= m[d, d]
            if not (lo.any() or hi.any()):
                break
            x[lo] += m[d]
            x[hi] -= m[d]
    return x


def sheared(system, offdiag_frac=(0.25, -0.2, 0.3)):
    """The same molecules in a TRICLINIC cell of the same volume: box vectors a = (L_x, 0, 0), b = (f0 L_x, L_y, 0),
    c = (f1 L_x, f2 L_y, L_z) (within the reference's limits |b_x|, |c_x| <= a_x / 2, |c_y| <= b_y / 2), every atom carried
    along by the shear (fractional coordinates kept), then put into the brick the grid covers.  Molecules are deformed a little;
    the nonbonded parity tests do not care."""
    b = system.box.astype(np.float64)
    off = np.array([offdiag_frac[0] * b[0], offdiag_frac[1] * b[0], offdiag_frac[2] * b[1]], np.float32)
    m = np.array([[b[0], 0, 0], [off[0], b[1], 0], [off[1], off[2], b[2]]], np.float64)
    frac = system.x.astype(np.float64) / b
    x = put_atoms_in_triclinic_box((frac @ m).astype(np.float32), m.astype(np.float32))
    out = System(x, system.box.copy(), system.types, system.q, system.nbfp, system.excl_off, system.excl_idx, system.mol_id,
                 system.name + "_triclinic")
    out.box_offdiag = off
    return out


def perturbed_water(name="water_3k", nmol=20, seed=7, q_scale_b=0.25):
    """A water box with `nmol` perturbed molecules for the free-energy tests: in state B their charges are scaled by q_scale_b and
    the oxygen's Lennard-Jones interaction is switched off (type 1: a disappearing particle, what soft-core exists for).
    Returns (system, perturbed[n] bool, typeA, typeB, qA, qB, types_masked, q_masked) -- the masked arrays are what the cluster-pair
    path gets (nbnxn_atomdata_mask_fep, nbnxm/atomdata.cpp: perturbed atoms carry zero charge and no LJ there)."""
    s = name if isinstance(name, System) else named(name)
    rng = np.random.Generator(np.random.PCG64(seed))
    mols = rng.choice(s.n // 3, nmol, replace=False)
    pert = np.zeros(s.n, bool)
    for m in mols:
        pert[3 * m:3 * m + 3] = True
    qA, qB = s.q.copy(), s.q.copy()
    qB[pert] *= np.float32(q_scale_b)
    tA, tB = s.types.copy(), s.types.copy()
    tB[pert] = 1
    tm, qm = s.types.copy(), s.q.copy()
    tm[pert] = 1
    qm[pert] = 0
    return s, pert, tA, tB, qA, qB, tm, qm


FEP_CASES = {  # name: parameters of the reference's soft-core (t_lambda) and the two lambdas
    "sc1": dict(lambda_coul=0.3, lambda_vdw=0.6, sc_alpha=0.5, sc_power=1, sc_sigma=0.3, sc_sigma_min=0.3, sc_coul=False),
    "nosc": dict(lambda_coul=0.4, lambda_vdw=0.4, sc_alpha=0.0, sc_power=1, sc_sigma=0.3, sc_sigma_min=0.3, sc_coul=False),
    "sc2coul": dict(lambda_coul=0.5, lambda_vdw=0.5, sc_alpha=0.7, sc_power=2, sc_sigma=0.3, sc_sigma_min=0.25, sc_coul=True),
    "sc1coul": dict(lambda_coul=0.2, lambda_vdw=0.9, sc_alpha=0.5, sc_power=1, sc_sigma=0.3, sc_sigma_min=0.3, sc_coul=True),
}


FEP_TWIN = (  # tag, LJ-PME rule, rvdw_switch, soft-core case -- the rvdw = 0.8 < rcoulomb = 0.9 cases of the free-energy tests
    ("cut_sc1", 0, 0.0, "sc1"), ("cut_sc2coul", 0, 0.0, "sc2coul"), ("pswitch_nosc", 0, 0.7, "nosc"), ("ljpme_geom_sc1coul", 1, 0.0, "sc1coul"))


def nbfp_two_lj_types(sigma_h=0.12, eps_h=0.19):
    """A second nonbonded-parameter table for the water boxes in which the hydrogens carry Lennard-Jones parameters too
    (a TIP-like sigma / epsilon pair) and the O-H cross term follows Lorentz-Berthelot: two LJ types whose geometric and
    Lorentz-Berthelot combinations differ, for the tests of the combination-rule dependent paths (type table, LJ-PME)."""
    c6 = {0: C6_O, 1: 4 * eps_h * sigma_h ** 6}
    c12 = {0: C12_O, 1: 4 * eps_h * sigma_h ** 12}
    sig = {t: (c12[t] / c6[t]) ** (1.0 / 6.0) for t in (0, 1)}
    eps = {t: c6[t] ** 2 / (4 * c12[t]) for t in (0, 1)}
    nbfp = np.zeros((2, 2, 2), np.float32)
    for a in (0, 1):
        for b in (0, 1):
            s_ab, e_ab = 0.5 * (sig[a] + sig[b]), (eps[a] * eps[b]) ** 0.5
            nbfp[a, b] = (6.0 * 4 * e_ab * s_ab ** 6, 12.0 * 4 * e_ab * s_ab ** 12)
    return nbfp


def perturbed_water_ljpme(nmol=20, seed=7, q_scale_b=0.25):
    """perturbed_water for the LJ-PME tests of the free-energy kernel: the two-LJ-type table (so that the geometric and the
    Lorentz-Berthelot grid rule differ) extended by a third type WITHOUT Lennard-Jones, which the perturbed atoms -- oxygens and
    hydrogens, both with LJ in state A -- take in state B (disappearing particles: the soft-core is active, and their grid C6
    vanishes in state B) and in the masked atom data of the cluster-pair path.  Returns (system, perturbed, typeA, typeB, qA, qB,
    types_masked, q_masked, nbfp[3,3,2])."""
    s, pert, tA, tB, qA, qB, tm, qm = perturbed_water(nmol=nmol, seed=seed, q_scale_b=q_scale_b)
    nbfp = np.zeros((3, 3, 2), np.float32)
    nbfp[:2, :2] = nbfp_two_lj_types()
    tB, tm = tA.copy(), tA.copy()
    tB[pert] = 2
    tm[pert] = 2
    return s, pert, tA, tB, qA, qB, tm, qm, nbfp


NAMED = {
    "water_3k": (10, 10, 10),
    "water_24k": (20, 20, 20),      # BASELINE.json configs[1]
    "water_96k": (40, 40, 20),      # configs[2]
    "water_192k": (40, 40, 40),
    "water_1M": (70, 70, 70),       # configs[3]: 1.029 M atoms, 21.75 nm
    "water_1.5M": (80, 80, 80),     # configs[4] per-GPU tile: 1.536 M atoms
}


# The reference's own benchmark water: BenchmarkSystem(S) stacks coordinates1000 (1000 SPC/E molecules, liquid structure,
# 3.10736 nm box) 2 x 2 x ... times (nbnxm/benchmark/bench_system.cpp:90-151); tiles per dimension here
REF_NAMED = {
    "ref_water_3k": (1, 1, 1),
    "ref_water_24k": (2, 2, 2),     # BenchmarkSystem(8):  BASELINE.json configs[1] as SURVEY 8(d) defines it
    "ref_water_96k": (4, 4, 2),     # BenchmarkSystem(32): configs[2]
    "ref_water_192k": (4, 4, 4),
    "ref_water_1M": (7, 7, 7),      # configs[3]: 1.029 M atoms (the reference's tiler only does powers of two)
}


def ref_water_box(tx, ty, tz):
    """generateCoordinates + BenchmarkSystem (bench_system.cpp:90-195): the base tile shifted by whole boxes, x outermost and z
    innermost, put in the box; O type 0 / H type 1, SPC/E charges, intramolecular exclusions."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ref_water_1000.npz"))
    base, edge = d["x"].astype(np.float32), np.float32(d["box_edge"])
    sh = np.stack(np.meshgrid(np.arange(tx), np.arange(ty), np.arange(tz), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    x = (base[None, :, :] + (sh * edge)[:, None, :]).reshape(-1, 3).astype(np.float32)
    box = (np.array([tx, ty, tz], np.float32) * edge).astype(np.float32)
    x = np.where(x >= box, x - box, x)
    x = np.where(x < 0, x + box, x).astype(np.float32)  # put_atoms_in_box, bench_system.cpp:157
    n = x.shape[0]
    nmol = n // 3
    types = np.tile(np.array([0, 1, 1], np.int32), nmol)
    q = np.tile(np.array([Q_O, Q_H, Q_H], np.float32), nmol)
    nbfp = np.zeros((2, 2, 2), np.float32)
    nbfp[0, 0] = (6.0 * C6_O, 12.0 * C12_O)
    first = (np.arange(n) // 3) * 3
    excl_idx = (first[:, None] + np.arange(3)[None, :]).astype(np.int32).ravel()
    excl_off = (np.arange(n + 1) * 3).astype(np.int32)
    return System(x, box, types, q, nbfp, excl_off, excl_idx, (np.arange(n) // 3).astype(np.int32), "ref_water_%dx%dx%d" % (tx, ty, tz))


def tiles_of(name):
    """(generator, (nx, ny, nz)) of a named workload: molecules per dimension (lattice water) or tiles (reference water)."""
    if name in REF_NAMED:
        return ref_water_box, REF_NAMED[name]
    return water_box, NAMED[name]


def named(name, seed=20261017):
    if name in REF_NAMED:
        return ref_water_box(*REF_NAMED[name])
    return water_box(*NAMED[name], seed=seed)


def argon12():
    """The nblib argon sample: api/nblib/samples/argon-forces-integration.cpp:52-57,86-88,103."""
    x = np.array([[0.794, 1.439, 0.610], [1.397, 0.673, 1.916], [0.659, 1.080, 0.573],
                  [1.105, 0.090, 3.431], [1.741, 1.291, 3.432], [1.936, 1.441, 5.873],
                  [0.960, 2.246, 1.659], [0.382, 3.023, 2.793], [0.053, 4.857, 4.242],
                  [2.655, 5.057, 2.211], [4.114, 0.737, 0.614], [5.977, 5.104, 5.217]], np.float32)
    n = 12
    nbfp = np.array([[[6 * 0.0062647225, 12 * 9.847044e-06]]], np.float32)
    return System(x, np.array([6.05449] * 3, np.float32), np.zeros(n, np.int32), np.zeros(n, np.float32), nbfp,
                  np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32),
                  "argon12")


def ewald_beta(rc, rtol=1e-5):
    """calc_ewaldcoeff_q (src/gromacs/ewald/ewald_utils.cpp:46-74): bisection on erfc(beta*rc) = rtol."""
    from math import erfc
    beta = 5.0
    i = 0
    while erfc(beta * rc) > rtol:
        i += 1
        beta *= 2
    n = i + 60
    low, high = 0.0, beta
    for _ in range(n):
        beta = (low + high) / 2
        if erfc(beta * rc) > rtol:
            low = beta
        else:
            high = beta
    return beta


def ewald_beta_lj(rc, rtol=1e-3):
    """calc_ewaldcoeff_lj (src/gromacs/ewald/ewald_utils.cpp:76-115): bisection on exp(-x^2)(1 + x^2 + x^4/2) = rtol, x = beta*rc
    (ewald-rtol-lj defaults to 1e-3)."""
    from math import exp

    def fn(beta):
        x2 = (beta * rc) ** 2
        return exp(-x2) * (1 + x2 + x2 * x2 / 2.0)

    beta = 5.0
    i = 0
    while True:
        i += 1
        beta *= 2
        if not fn(beta) > rtol:
            break
    low, high = 0.0, beta
    for _ in range(i + 60):
        beta = (low + high) / 2
        if fn(beta) > rtol:
            low = beta
        else:
            high = beta
    return beta


def rf_constants(rc, eps_rf=0.0, eps_r=1.0):
    """calc_rffac (src/gromacs/mdlib/rf_util.cpp); eps_rf = 0 means infinity as in
    benchmark/bench_setup.cpp:152-155 (k_rf = 0.5/rc^3, c_rf = 1/rc + k_rf rc^2)."""
    if eps_rf == 0:
        k = 1.0 / (2 * rc ** 3)
    else:
        k = (eps_rf - eps_r) / ((2 * eps_rf + eps_r) * rc ** 3)
    return k, 1.0 / rc + k * rc * rc


def spc_methanol():
    """The nblib SPC-water + methanol test system: api/nblib/tests/testsystems.cpp:56-86 (parameters),
    :107-153 (charges, exclusions), :313-326 (coordinates, box 3.01).  nblib combines C6/C12 with the
    geometric rule (api/nblib/gmxsetup.cpp:143-166)."""
    x = np.array([[1.970, 1.460, 1.209], [1.978, 1.415, 1.082], [1.905, 1.460, 1.030],
                  [1.555, 1.511, 0.703], [1.498, 1.495, 0.784], [1.496, 1.521, 0.623]], np.float32)
    # types: 0 Ow, 1 H, 2 OMet, 3 CMet
    c6 = np.array([0.0026173456, 0.0, 0.0022619536, 0.0088755241])
    c12 = np.array([2.634129e-06, 0.0, 1.505529e-06, 2.0852922e-05])
    nbfp = np.zeros((4, 4, 2), np.float32)
    nbfp[..., 0] = 6.0 * np.sqrt(np.outer(c6, c6))
    nbfp[..., 1] = 12.0 * np.sqrt(np.outer(c12, c12))
    types = np.array([3, 2, 1, 0, 1, 1], np.int32)
    q = np.array([0.176, -0.574, 0.398, -0.82, 0.41, 0.41], np.float32)
    n = 6
    first = (np.arange(n) // 3) * 3
    excl_idx = (first[:, None] + np.arange(3)[None, :]).astype(np.int32).ravel()
    excl_off = (np.arange(n + 1) * 3).astype(np.int32)
    return System(x, np.array([3.01] * 3, np.float32), types, q, nbfp, excl_off, excl_idx,
                  (np.arange(n) // 3).astype(np.int32), "spc_methanol")


def bonded_chains(nchains=40, length=24, box=(3.1, 2.9, 3.3), seed=13, box_matrix=None):
    """Flexible chains for the listed-interaction ("bonded") tests: `nchains` random-walk chains of `length` atoms (bond length
    ~0.15 nm, bend and torsion angles well away from 0 / 180 degrees), wrapped atom by atom into a rectangular periodic box so
    that many interactions straddle a face (`box_matrix`: the same chains wrapped into that triclinic cell instead; `box` is then
    its diagonal).  Returns a dict: x[n,3], q[n], box[3], and per interaction type of the reference's
    GPU bonded module (listed_forces/gpubonded.h:84-85) `iatoms` = {parameter index, atoms...} rows and `params` = rows of 6
    floats (the t_iparams fields the type reads):
      bonds {r0, k} | angles {theta0 deg, k} | urey_bradley {theta0, ktheta, r13, kUB} | pdihs, pidihs {phi0 deg, k, multiplicity}
      | rbdihs {C0..C5} | idihs {xi0 deg, k} | lj14 {c6, c12}."""
    rng = np.random.Generator(np.random.PCG64(seed))
    box = np.asarray(box, np.float32)
    M = np.diag(box).astype(np.float64) if box_matrix is None else np.asarray(box_matrix, np.float64)
    box = np.diag(M).astype(np.float32)

    def min_image_norm(d):
        for dim in (2, 1, 0):
            d = d - np.rint(d[:, dim] / M[dim, dim])[:, None] * M[dim]
        return np.sqrt((d * d).sum(1))

    xs = []  # finished chains; no two atoms that are not bonded neighbours closer than 0.2 nm (a liquid, not overlapping walks)

    def clear_of(q, others):
        return len(others) == 0 or min_image_norm(np.asarray(others) - q).min() >= 0.2

    while len(xs) < nchains:
        done = np.concatenate(xs) if xs else np.zeros((0, 3))
        p = [rng.uniform(0, box)]
        if not clear_of(p[0], done):
            continue
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        while len(p) < length:
            for _ in range(40):
                # turn the direction by 50-80 degrees around a random axis: bend angles 100-130 degrees, torsions anywhere
                axis = np.cross(d, rng.normal(size=3))
                axis /= np.linalg.norm(axis)
                ang = np.deg2rad(rng.uniform(50, 80))
                dn = d * np.cos(ang) + np.cross(axis, d) * np.sin(ang) + axis * np.dot(axis, d) * (1 - np.cos(ang))
                q = p[-1] + dn * rng.uniform(0.13, 0.17)
                if clear_of(q, done) and clear_of(q, p[:-3]):
                    p.append(q)
                    d = dn
                    break
            else:
                break  # dead end: start this chain again
        if len(p) == length:
            xs.append(np.array(p))
    x = put_atoms_in_triclinic_box(np.concatenate(xs).astype(np.float32), M.astype(np.float32))
    n = nchains * length
    q = rng.uniform(-0.6, 0.6, n).astype(np.float32)
    first = np.arange(nchains) * length

    def rows(nat, ntypes, step=1):
        out = []
        for c in first:
            for a in range(0, length - nat + 1, step):
                out.append([rng.integers(ntypes)] + [c + a + k for k in range(nat)])
        return np.array(out, np.int32)

    def p6(*cols):
        m = np.zeros((len(cols[0]), 6), np.float32)
        for k, c in enumerate(cols):
            m[:, k] = c
        return m

    s = dict(x=x, q=q, box=box, n=n)
    s["bonds"] = dict(iatoms=rows(2, 3), params=p6([0.14, 0.15, 0.16], [2.5e5, 3.0e5, 2.0e5]))
    s["angles"] = dict(iatoms=rows(3, 3), params=p6([109.5, 120.0, 114.0], [400.0, 520.0, 350.0]))
    s["urey_bradley"] = dict(iatoms=rows(3, 2, step=2), params=p6([110.0, 118.0], [300.0, 420.0], [0.24, 0.26], [2.0e4, 3.0e4]))
    s["pdihs"] = dict(iatoms=rows(4, 4), params=p6([0.0, 180.0, 60.0, 35.0], [4.0, 6.5, 1.2, 9.0], [3, 2, 1, 4]))
    s["rbdihs"] = dict(iatoms=rows(4, 2, step=2), params=np.array([[9.28, 12.16, -13.12, -3.06, 26.24, -31.5],
                                                                     [2.0, -1.5, 0.7, 3.3, -0.4, 0.9]], np.float32))
    imp = rows(4, 2, step=3)
    imp[:, 1:] = imp[:, [2, 1, 3, 4]]  # centre atom first, as an improper lists them
    s["idihs"] = dict(iatoms=imp, params=p6([35.3, -20.0], [167.0, 334.0]))
    pimp = rows(4, 2, step=5)
    s["pidihs"] = dict(iatoms=pimp, params=p6([180.0, 0.0], [4.6, 43.9], [2, 2]))
    s["lj14"] = dict(iatoms=rows(4, 3)[:, [0, 1, 4]].copy(), params=p6([2.3e-3, 1.1e-3, 0.0], [2.5e-6, 9.0e-7, 4.0e-7]))
    return s
