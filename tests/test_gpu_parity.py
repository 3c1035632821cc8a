"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): in-range pair set bit-exact; per-atom forces <= 1e-5 relative RMS;
shift forces / virial <= 1e-5 relative; energies: see test_energies for the conditioning note.
"""
import numpy as np
import pytest

import gmxapi_b200 as g
from gmxapi_b200 import lib as nb
from oracle import oracle

pytestmark = pytest.mark.gpu

RC = 0.9
ENERGY_TOL = 2e-5  # relative, against the oracle's double-precision sums (the reference's own FP32 sums sit 1e-5 .. 2e-4 from those)
FORCE_TOL = 1e-5  # relative RMS, north_star
VIRIAL_TOL = 1e-5


def relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def make(s, coulomb, rc=RC, energy=True, rlo=0.0, rli=0.0):
    opt = g.NBKernelOptions(pairlistCutoff=rc, coulombType=coulomb, computeVirialAndEnergy=energy, rlistOuter=rlo,
                            rlistInner=rli)
    return g.ForceCalculator(g.SimulationState.from_system(s), opt)


def oracle_kwargs(coulomb, rc=RC):
    if coulomb == g.CoulombType.Pme:
        return dict(eeltype=oracle.EEL_EWALD, beta=float(np.float32(g.systems.ewald_beta(rc))))
    k, c = g.systems.rf_constants(rc, eps_rf=1.0)
    if coulomb == g.CoulombType.Cutoff:
        return dict(eeltype=oracle.EEL_CUT, k_rf=k, c_rf=c)
    return dict(eeltype=oracle.EEL_RF, k_rf=k, c_rf=c)


def test_argon_golden(built):
    """api/nblib/tests/nbkernelsystem.cpp:187-202 ArgonForcesAreCorrect against
    refdata/NBlibTest_ArgonForcesAreCorrect.xml (copied as tests/golden/argon12_forces.json)."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "argon12_forces.json")))
    s = g.systems.argon12()
    fc = make(s, g.CoulombType.Cutoff, rc=1.0)
    f = fc.compute()
    ref = np.array(gold["forces"], np.float64)
    # the reference's own tolerance is 200 ULP float relative (api/nblib/tests/testhelpers.h:73-77)
    assert np.allclose(f, ref, rtol=2e-5, atol=1e-8)
    assert fc.nb.pair_count(1.0) == 1


@pytest.mark.parametrize("name", ["water_3k", "water_24k"])
def test_grid_order_matches_oracle(built, name):
    s = g.systems.named(name)
    fc = make(s, g.CoulombType.ReactionField, energy=False)
    go = oracle.put_on_grid(s.x, s.box)
    st = fc.nb.stats()
    assert (st["ncx"], st["ncy"], st["natoms_padded"]) == (go["ncx"], go["ncy"], go["npad"])
    assert np.array_equal(fc.nb.grid_order(), go["atom_index"])


@pytest.mark.parametrize("name,coulomb", [("water_3k", g.CoulombType.Pme), ("water_3k", g.CoulombType.ReactionField),
                                          ("water_24k", g.CoulombType.Pme), ("water_24k", g.CoulombType.ReactionField),
                                          ("water_96k", g.CoulombType.Pme)])
def test_pairs_forces_energies(built, name, coulomb):
    s = g.systems.named(name)
    fc = make(s, coulomb)
    f = fc.compute()
    fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx,
                                              **oracle_kwargs(coulomb))
    # 1. pair set, bit-exact
    gp = oracle.canonical_pairs(fc.nb.pairs(RC))
    op = oracle.canonical_pairs(oracle.pair_set(s.x, s.box, RC, s.excl_off, s.excl_idx))
    assert len(gp) == len(op) == npairs
    assert np.array_equal(gp, op)
    # 2. forces
    assert relrms(f, fo) < FORCE_TOL
    # 3. shift forces (the central shift carries no virial: shift_vec = 0) and the virial built from them
    m = np.ones(45, bool)
    m[nb.CENTRAL] = False
    fs = fc.shiftForces.astype(np.float64)
    assert np.abs(fs[m] - fso[m]).max() <= VIRIAL_TOL * np.abs(fso[m]).max()
    vo = oracle.virial_from_fshift(s.box, fso)
    vg = oracle.virial_from_fshift(s.box, np.where(m[:, None], fs, 0.0))
    assert np.abs(vg - vo).max() <= VIRIAL_TOL * np.abs(vo).max()
    # 4. energies, against the oracle's DOUBLE-precision sums of the same single-precision pair terms.  The kernel keeps per-warp
    # single-precision partial sums and adds those in double, which holds 2e-5 of the total; the reference's own all-FP32 sums
    # sit 1e-5 .. 1.8e-4 from the double sum, hence the looser 2e-4 where the expected value is a reference output.
    elj, eel = fc.energies
    assert abs(elj - evo) <= ENERGY_TOL * abs(evo)
    assert abs(eel - eco) <= ENERGY_TOL * abs(eco)


DEFAULT_SHEAR = (0.25, -0.2, 0.3)


@pytest.mark.parametrize("coulomb,rlo,rli,shear", [(g.CoulombType.Pme, 0.0, 0.0, DEFAULT_SHEAR), (g.CoulombType.ReactionField, 0.0, 0.0, DEFAULT_SHEAR),
                                                   (g.CoulombType.Pme, 1.05, 0.95, DEFAULT_SHEAR),
                                                   # at the limits of check_box: 1169 pairs sit two box vectors away along x
                                                   (g.CoulombType.Pme, 0.0, 0.0, (0.5, 0.5, -0.5))])
def test_triclinic_cell(built, coulomb, rlo, rli, shear):
    """A triclinic cell (b200nb_set_box_triclinic: the 3 k water box sheared, box[YY][XX] = 0.25 L, box[ZZ][XX] = -0.2 L,
    box[ZZ][YY] = 0.3 L): shift vectors k a + l b + m c, x-shift range 2.  Pair set bit-exact, forces, shift forces / virial and
    energies against the triclinic mode of the oracle, and the forces against the reference's own output on the same cell
    (tests/golden/ref_water_3k_triclinic_*.npz); once with a list buffer and dynamic pruning."""
    import os
    s = g.systems.sheared(g.systems.named("water_3k"), shear)
    fc = make(s, coulomb, rlo=rlo, rli=rli)
    f = fc.compute()
    oracle.set_triclinic(s.box_offdiag)
    try:
        fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, **oracle_kwargs(coulomb))
        op = oracle.canonical_pairs(oracle.pair_set(s.x, s.box, RC, s.excl_off, s.excl_idx))
        sv = oracle.shift_vectors(s.box).astype(np.float64)
    finally:
        oracle.set_triclinic(None)
    gp = oracle.canonical_pairs(fc.nb.pairs(RC))
    assert len(gp) == len(op) == npairs and np.array_equal(gp, op)
    if shear != DEFAULT_SHEAR:
        assert np.any(np.isin(gp & 63, [0, 4, 5, 9, 10, 14, 15, 19, 20]))  # shifts with |t_x| = 2 and index <= CENTRAL are in use
    assert relrms(f, fo) < FORCE_TOL
    m = np.ones(45, bool)
    m[nb.CENTRAL] = False
    fs = fc.shiftForces.astype(np.float64)
    assert np.abs(fs[m] - fso[m]).max() <= VIRIAL_TOL * np.abs(fso[m]).max()
    vo, vg = -0.5 * sv.T @ fso, -0.5 * sv.T @ np.where(m[:, None], fs, 0.0)
    assert np.abs(vo).max() > 0 and np.abs(vg - vo).max() <= VIRIAL_TOL * np.abs(vo).max()
    elj, eel = fc.energies
    assert abs(elj - evo) <= ENERGY_TOL * abs(evo) and abs(eel - eco) <= ENERGY_TOL * abs(eco)
    if coulomb == g.CoulombType.Pme and shear == DEFAULT_SHEAR:  # (the reaction-field fixture has epsilon_rf = infinity, this run 1)
        gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_water_3k_triclinic_ewald.npz"))
        assert len(gp) == int(gd["npairs"]) and relrms(f, gd["f"].astype(np.float64)) < FORCE_TOL
    # a list radius the cell cannot hold is refused (max_cutoff2), as is a box matrix outside the reference's limits
    with pytest.raises(nb.B200NBError):
        make(s, coulomb, rc=1.6)
    with pytest.raises(nb.B200NBError):
        fc.nb.set_box(np.array([[3.0, 0, 0], [1.6, 3.0, 0], [0, 0, 3.0]], np.float32))
    fc.nb.close()


@pytest.mark.parametrize("flavour,mod,rvdw,rsw", [("twin", g.VdwModifier.PotentialShift, 0.8, 0.0),
                                                  ("fswitch", g.VdwModifier.ForceSwitch, 0.9, 0.75),
                                                  ("pswitch", g.VdwModifier.PotentialSwitch, 0.9, 0.75),
                                                  ("fswitch_twin", g.VdwModifier.ForceSwitch, 0.8, 0.65)])
def test_vdw_flavours(built, flavour, mod, rvdw, rsw):
    """LJ force switch, potential switch and the twin-range VdW cut-off (the general kernels, b200nb_set_vdw) against the
    oracle and against the committed outputs of the reference's SIMD kernels; once with the water charges, once with all
    charges zero so that the Lennard-Jones arithmetic is what the force tolerance measures."""
    import os
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_water_3k_vdw_%s.npz" % flavour))
    s = g.systems.named("water_3k")
    beta = float(np.float32(g.systems.ewald_beta(RC)))
    for tag, q in (("", s.q), ("_lj", np.zeros_like(s.q))):
        for energy in (True, False):
            opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=energy,
                                    vdwModifier=mod, vdwCutoff=rvdw, vdwSwitch=rsw)
            fc = g.ForceCalculator(g.SimulationState(s.x, s.box, s.types, q, s.nbfp, s.excl_off, s.excl_idx), opt)
            f = fc.compute()
            fo, fso, evo, eco, _ = oracle.forces(s.x, s.box, q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx,
                                                 eeltype=oracle.EEL_EWALD, beta=beta, vdw_modifier=mod.value,
                                                 rvdw=rvdw, rvdw_switch=rsw)
            assert relrms(f, fo) < FORCE_TOL
            assert relrms(f, gd["f" + tag].astype(np.float64)) < FORCE_TOL
            if energy:
                elj, eel = fc.energies
                # the switched LJ energy of this box is a near-cancelling sum (8 kJ/mol against 113 with the plain cut-off): the
                # bar is ENERGY_TOL of the larger of the two, i.e. the same absolute error as for the plain potential
                assert abs(elj - evo) <= ENERGY_TOL * max(abs(evo), 113.0)
                assert abs(elj - float(gd["e_lj" + tag])) <= 2e-4 * abs(float(gd["e_lj" + tag]))
                if tag == "":
                    assert abs(eel - eco) <= ENERGY_TOL * abs(eco)
                m = np.ones(45, bool)
                m[nb.CENTRAL] = False
                fs = fc.shiftForces.astype(np.float64)
                assert np.abs(fs[m] - fso[m]).max() <= VIRIAL_TOL * np.abs(fso[m]).max()
            fc.nb.close()
    # rvdw < rcoulomb without Ewald is refused, as in the reference's Verlet scheme
    with pytest.raises(nb.B200NBError):
        g.ForceCalculator(g.SimulationState.from_system(s),
                          g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.ReactionField, vdwCutoff=0.8))


@pytest.mark.parametrize("rule,ljpme", [("geom", g.LjPme.Geometric), ("lb", g.LjPme.LorentzBerthelot)])
def test_ljpme_grid_correction(built, rule, ljpme):
    """LJ-PME real-space kernels (grid part of the dispersion subtracted; geometric and Lorentz-Berthelot grid rule) against
    the oracle and the committed outputs of the reference kernels, with the water charges and with all charges zero; the
    system has LJ on the hydrogens too, so the two rules differ (4e-4 of the forces) and the type-table kernels run."""
    import os
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_water_3k_ljpme_%s.npz" % rule))
    s = g.systems.named("water_3k")
    nbfp = g.systems.nbfp_two_lj_types()
    beta = float(np.float32(g.systems.ewald_beta(RC)))
    bl = float(gd["ewaldcoeff_lj"])
    for tag, q in (("", s.q), ("_lj", np.zeros_like(s.q))):
        for energy in (True, False):
            opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=energy, ljPme=ljpme)
            fc = g.ForceCalculator(g.SimulationState(s.x, s.box, s.types, q, nbfp, s.excl_off, s.excl_idx), opt)
            assert not fc.nb.stats()["comb_geometric"]  # Lorentz-Berthelot cross terms: the type table
            f = fc.compute()
            fo, fso, evo, eco, _ = oracle.forces(s.x, s.box, q, s.types, nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD,
                                                 beta=beta, ljpme=ljpme.value, ewaldcoeff_lj=bl)
            assert relrms(f, fo) < FORCE_TOL
            assert relrms(f, gd["f" + tag].astype(np.float64)) < FORCE_TOL
            if energy:
                elj, eel = fc.energies
                assert abs(elj - evo) <= ENERGY_TOL * abs(evo)
                assert abs(elj - float(gd["e_lj" + tag])) <= 2e-4 * abs(float(gd["e_lj" + tag]))
                if tag == "":
                    assert abs(eel - eco) <= ENERGY_TOL * abs(eco)
                m = np.ones(45, bool)
                m[nb.CENTRAL] = False
                fs = fc.shiftForces.astype(np.float64)
                assert np.abs(fs[m] - fso[m]).max() <= VIRIAL_TOL * np.abs(fso[m]).max()
            fc.nb.close()
    # geometric LJ parameters (plain water) with LJ-PME: the library switches to the type-table kernels by itself
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True, ljPme=ljpme)
    fc = g.ForceCalculator(g.SimulationState.from_system(s), opt)
    assert fc.nb.stats()["comb_geometric"]
    f = fc.compute()
    fo, _, evo, _, _ = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD, beta=beta,
                                     ljpme=ljpme.value, ewaldcoeff_lj=float(np.float32(g.systems.ewald_beta_lj(RC))))
    assert relrms(f, fo) < FORCE_TOL and abs(fc.energies[0] - evo) <= ENERGY_TOL * abs(evo)
    # LJ-PME with a switch modifier is refused, as in the reference (vdwtype PME implies potential shift)
    with pytest.raises(g.nblib.InputException):
        g.ForceCalculator(g.SimulationState.from_system(s), g.NBKernelOptions(pairlistCutoff=RC, ljPme=ljpme, vdwSwitch=0.7,
                                                                               vdwModifier=g.VdwModifier.ForceSwitch))


def test_tile_list_is_exact(built):
    """The device list holds exactly the cluster pairs with >= 1 atom pair inside rlist (what the reference list
    converges to after pruning), in the half-list convention."""
    s = g.systems.named("water_24k")
    fc = make(s, g.CoulombType.ReactionField, energy=False)
    go = oracle.put_on_grid(s.x, s.box)
    to = oracle.tile_list(s.x, s.box, RC, go["slot_of_atom"])
    tg = fc.nb.tiles()
    key = lambda t: np.sort((t[:, 0].astype(np.int64) << 40) | (t[:, 1].astype(np.int64) << 32) | t[:, 2].astype(np.int64))
    assert len(tg) == len(to)
    assert np.array_equal(key(tg), key(to))


def test_dynamic_pruning(built):
    """Outer list at 1.05, inner at 0.95: the pruned list equals the oracle's prune of the outer list, results with
    and without pruning agree, and a rolling prune after moving the atoms keeps the pair set exact."""
    s = g.systems.named("water_24k")
    fc = make(s, g.CoulombType.Pme, rlo=1.05, rli=0.95, energy=False)
    go = oracle.put_on_grid(s.x, s.box)
    outer = fc.nb.tiles(outer=True)
    inner = fc.nb.tiles()
    keep = oracle.prune_tiles(outer, go["atom_index"], s.x, s.box, 0.95)
    key = lambda t: np.sort((t[:, 0].astype(np.int64) << 40) | (t[:, 1].astype(np.int64) << 32) | t[:, 2].astype(np.int64))
    assert 0 < keep.sum() < len(outer)
    assert np.array_equal(key(inner), key(outer[keep]))
    f1 = fc.compute()
    fo = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, energy=False,
                       **oracle_kwargs(g.CoulombType.Pme))[0]
    assert relrms(f1, fo) < FORCE_TOL
    # move atoms by up to 0.02 nm (well inside the 0.05 nm half-buffer), rolling prune in 4 parts, recompute
    rng = np.random.Generator(np.random.PCG64(7))
    x2 = (s.x + rng.uniform(-0.02, 0.02, s.x.shape)).astype(np.float32)
    fc.nb.set_x(x2)
    for part in range(4):
        fc.nb.launch_prune(-1, part, 4)
    f2 = fc.compute(x2)
    fo2, _, _, _, np2 = oracle.forces(x2, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, energy=False,
                                      **oracle_kwargs(g.CoulombType.Pme))
    assert relrms(f2, fo2) < FORCE_TOL
    gp = oracle.canonical_pairs(fc.nb.pairs(RC))
    op = oracle.canonical_pairs(oracle.pair_set(x2, s.box, RC, s.excl_off, s.excl_idx))
    assert np.array_equal(gp, op)


def test_type_table_path(built):
    """LJ parameters that do NOT follow the geometric rule take the type-table kernel (nbnxm atomdata.cpp:462-525)."""
    s = g.systems.named("water_3k")
    s.nbfp = s.nbfp.copy()
    s.nbfp[0, 1] = s.nbfp[1, 0] = (6 * 0.001, 12 * 1e-6)  # O-H cross term, not sqrt(c_OO * c_HH) = 0
    fc = make(s, g.CoulombType.Pme)
    assert fc.nb.stats()["comb_geometric"] == 0
    f = fc.compute()
    fo = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx,
                       **oracle_kwargs(g.CoulombType.Pme))[0]
    assert relrms(f, fo) < FORCE_TOL


def test_edge_cases(built):
    """Ragged / tiny inputs: a single cluster, a column with one atom, atoms exactly on the box edge."""
    rng = np.random.Generator(np.random.PCG64(3))
    for n in (2, 9, 65, 200):
        box = np.array([3.0, 3.5, 4.0], np.float32)
        x = (rng.uniform(0, 1, (n, 3)) * box).astype(np.float32)
        x[0] = 0.0
        from gmxapi_b200.systems import System
        nbfp = np.array([[[6 * 0.0026, 12 * 2.6e-6]]], np.float32)
        q = rng.uniform(-0.5, 0.5, n).astype(np.float32)
        s = System(x, box, np.zeros(n, np.int32), q, nbfp, np.arange(n + 1, dtype=np.int32),
                   np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32), "rand%d" % n)
        fc = make(s, g.CoulombType.ReactionField)
        f = fc.compute()
        fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx,
                                                  **oracle_kwargs(g.CoulombType.ReactionField))
        gp = oracle.canonical_pairs(fc.nb.pairs(RC))
        op = oracle.canonical_pairs(oracle.pair_set(s.x, s.box, RC, s.excl_off, s.excl_idx))
        assert np.array_equal(gp, op)
        if npairs:
            assert relrms(f, fo) < FORCE_TOL
        else:
            assert np.abs(f).max() == 0


def test_errors_are_loud(built):
    s = g.systems.named("water_3k")
    h = nb.NbnxmGpu(0)
    with pytest.raises(g.B200NBError):
        h.set_atoms(s.types, s.q)  # params first
    h.set_params(s.nbfp, RC)
    with pytest.raises(g.B200NBError):
        h.set_atoms(s.types + 5, s.q)  # type out of range
    h.set_atoms(s.types, s.q, s.excl_off, s.excl_idx)
    with pytest.raises(g.B200NBError):
        h.build_pairlist()  # no grid
    h.set_box([1.0, 1.0, 1.0])
    h.put_on_grid(s.x, [0, 0, 0], [1.0, 1.0, 1.0])
    with pytest.raises(g.B200NBError):
        h.build_pairlist()  # box < 2*rlist


def test_full_size_properties(built):
    """BASELINE.json's 1M-atom configuration through size-independent properties: Newton's third law
    (sum of forces = 0), pair count = oracle count, translation of all atoms by a lattice vector keeps the
    pair count."""
    s = g.systems.named("water_1M")
    fc = make(s, g.CoulombType.Pme, energy=False)
    f = fc.compute()
    assert np.abs(f.astype(np.float64).sum(0)).max() < 1e-3 * np.abs(f).mean() * np.sqrt(s.n)
    n_gpu = fc.nb.pair_count(RC)
    n_orc = len(oracle.pair_set(s.x, s.box, RC, s.excl_off, s.excl_idx))
    assert n_gpu == n_orc


def test_packed_list_and_device_step(built):
    """The force kernel consumes the pruned list re-packed per j-atom (PackedList): fewer lanes than the 8x8 cluster-pair
    list, same pair set (test_pairs_forces_energies extracts the pairs from the packed list).  The device-resident
    step (b200nb_step) and the host call (b200nb_compute) with pinned and with pageable buffers give the same forces."""
    import torch
    s = g.systems.named("water_24k")
    fc = make(s, g.CoulombType.Pme, energy=False)
    st = fc.nb.stats()
    assert 0 < st["ntiles_packed"] < 0.75 * st["ntiles_inner"]
    fo = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, energy=False,
                       **oracle_kwargs(g.CoulombType.Pme))[0]
    f_pageable = fc.compute(s.x.copy())
    assert relrms(f_pageable, fo) < FORCE_TOL
    x_pin = torch.from_numpy(s.x.copy()).pin_memory()
    f_pin = torch.zeros_like(x_pin).pin_memory()
    fc.compute(x_pin.numpy(), f_pin.numpy())
    assert relrms(f_pin.numpy(), fo) < FORCE_TOL
    x_dev = torch.from_numpy(s.x).cuda()
    f_dev = torch.zeros_like(x_dev)
    torch.cuda.synchronize()
    for _ in range(2):  # twice: the step clears its own outputs
        fc.nb.step(x_dev.data_ptr(), f_dev.data_ptr(), 0)
    fc.nb.synchronize()
    assert relrms(f_dev.cpu().numpy(), fo) < FORCE_TOL
    # misaligned atom-order buffers take the scalar copy path
    xm = torch.zeros(3 * s.n + 1, dtype=torch.float32, device="cuda")
    xm[1:] = x_dev.reshape(-1)
    fm = torch.zeros(3 * s.n + 1, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    fc.nb.step(xm.data_ptr() + 4, fm.data_ptr() + 4, 0)
    fc.nb.synchronize()
    assert relrms(fm[1:].reshape(-1, 3).cpu().numpy(), fo) < FORCE_TOL
