"""-m gpu: the perturbed-pair (free-energy) kernel, gmxapi_b200/csrc/fep.cu, through the C ABI, against the restatement of the
reference's gmx_nb_free_energy_kernel (oracle/nbnxm_oracle.c orc_fep_kernel, pinned to the reference kernel itself in
tests/test_oracle_cpu.py) and against the committed outputs of the reference kernel (tests/golden/ref_water_3k_fep_rf.npz).
Tolerances: forces 1e-5 relative RMS (north_star), energies and dV/dlambda 2e-5 of the largest of the four sums."""
import os

import numpy as np
import pytest

import gmxapi_b200 as g
from gmxapi_b200 import lib as nb
from oracle import oracle

pytestmark = pytest.mark.gpu
RC = 0.9


def relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


@pytest.mark.parametrize("eel", ["rf", "ewald"])
@pytest.mark.parametrize("case", ["sc1", "nosc", "sc2coul", "sc1coul"])
def test_fep_kernel_matches_oracle_and_reference(built, case, eel):
    """20 perturbed water molecules in the 3 k box, reaction field.  The cluster-pair path runs on the MASKED atom data (perturbed atoms
    without charge and LJ, as nbnxn_atomdata_mask_fep leaves them), the free-energy kernel on the perturbed pair list; their sum is the
    force field of the lambda state.  Checked: the sum against oracle(masked system) + oracle(FEP list); the FEP part alone against
    the reference kernel's committed output; Vc, Vv, dV/dlambda."""
    S = g.systems
    s, pert, tA, tB, qA, qB, tm, qm = S.perturbed_water()
    kw = S.FEP_CASES[case]
    ewald = eel == "ewald"
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme if ewald else g.CoulombType.ReactionField,
                            computeVirialAndEnergy=True, ewaldPotentialShift=True)  # sh_ewald as mdrun sets it (and the fixture)
    masked = g.SimulationState(s.x, s.box, tm, qm, s.nbfp, s.excl_off, s.excl_idx)
    fc = g.ForceCalculator(masked, opt)
    h = fc.nb
    lst = oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx)
    h.fep_set_atoms(tA, tB, qA, qB)
    h.fep_upload_list(*lst)
    flags = nb.FLAG_ENERGY | nb.FLAG_VIRIAL
    # the cluster-pair kernels alone, then with the free-energy kernel between launch and read-back
    h.set_x(s.x)
    h.clear_outputs()
    h.launch_force(-1, flags)
    f_plain = h.get_f().copy()
    fs_plain = h.get_outputs()[0].astype(np.float64)
    h.set_x(s.x)
    h.clear_outputs()
    h.launch_force(-1, flags)
    h.fep_launch(**kw)
    f_sum = h.get_f().copy()
    fs_sum = h.get_outputs()[0].astype(np.float64)
    out4 = np.array(h.fep_outputs())
    # expected
    import math
    k, c = S.rf_constants(RC, eps_rf=1.0)
    beta = float(np.float32(S.ewald_beta(RC)))
    sv = oracle.shift_vectors(s.box)
    if ewald:
        fo_fep, fso_fep, o4 = oracle.fep_kernel(s.x, sv, s.nbfp, tA, tB, qA, qB, *lst, RC, ewaldcoeff=beta,
                                                sh_ewald=float(np.float32(math.erfc(beta * RC) / RC)), **kw)
        fo_plain = oracle.forces(s.x, s.box, qm, tm, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD, beta=beta)[0]
    else:
        fo_fep, fso_fep, o4 = oracle.fep_kernel(s.x, sv, s.nbfp, tA, tB, qA, qB, *lst, RC, k_rf=k, c_rf=c, **kw)
        fo_plain = oracle.forces(s.x, s.box, qm, tm, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_RF, k_rf=k, c_rf=c)[0]
    assert relrms(f_plain, fo_plain) < 1e-5
    f_fep = f_sum.astype(np.float64) - f_plain
    assert relrms(f_fep, fo_fep.astype(np.float64)) < 1e-5
    assert relrms(f_sum, fo_plain + fo_fep) < 1e-5
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_water_3k_fep_%s.npz" % eel))
    if ewald:
        assert abs(float(gd["sh_ewald"]) - math.erfc(beta * RC) / RC) < 1e-7 and abs(float(gd["beta"]) - beta) < 1e-6
    assert relrms(f_fep, gd["f_" + case].astype(np.float64)) < 1e-5
    m = np.ones(45, bool)
    m[nb.CENTRAL] = False
    fs_fep = fs_sum - fs_plain
    ref_fs = gd["fshift_" + case].astype(np.float64)
    assert np.abs(fs_fep[m] - ref_fs[m]).max() <= 2e-5 * np.abs(ref_fs[m]).max()
    o4r = gd["out4_" + case]
    assert np.abs(out4 - np.array(o4)).max() <= 2e-5 * np.abs(np.array(o4)).max()
    assert np.abs(out4 - o4r).max() <= 2e-5 * np.abs(o4r).max()
    # read and reset: nothing launched since
    assert h.fep_outputs() == (0.0, 0.0, 0.0, 0.0)
    fc.nb.close()


def _canonical(lst):
    """the pairs of a t_nblist as sorted keys (lower atom, higher atom, shift seen from the lower atom, flag)"""
    ii, sh, ji, jj, ex = [np.asarray(a) for a in lst]
    n = np.diff(ji)
    i, s, j = np.repeat(ii, n).astype(np.int64), np.repeat(sh, n).astype(np.int64), jj.astype(np.int64)
    swap = i > j
    a, b, s = np.where(swap, j, i), np.where(swap, i, j), np.where(swap, 44 - s, s)
    return np.sort((((a << 24) | b) << 8 | s) * 2 + ex.astype(np.int64))


@pytest.mark.parametrize("name,nmol,rlist", [("water_3k", 20, 0.9), ("water_3k", 20, 1.0), ("water_24k", 60, 1.0),
                                             ("water_3k_sheared", 30, 1.0), ("water_3k_sheared_hard", 30, 0.9)])
def test_fep_list_built_on_the_device(built, name, nmol, rlist):
    """b200nb_fep_build_list against the oracle's list (make_fep_list's pair set): the same pairs, shifts and exclusion flags --
    bit-exact as a set; a pair may be listed from the other atom, with the opposite shift -- and the kernel on the built list
    against the kernel on the uploaded one: forces, energies, dV/dlambda, virial."""
    S = g.systems
    if name.startswith("water_3k_sheared"):  # triclinic cells: a moderate shear and one at the limits of check_box
        base = S.sheared(S.named("water_3k"), (0.5, 0.5, -0.5) if name.endswith("hard") else (0.25, -0.2, 0.3))
    else:
        base = S.named(name)
    s, pert, tA, tB, qA, qB, tm, qm = S.perturbed_water(base, nmol=nmol)
    triclinic = bool(np.any(s.box_offdiag != 0))
    opt = g.NBKernelOptions(pairlistCutoff=RC, rlistOuter=rlist, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True)
    fc = g.ForceCalculator(g.SimulationState(s.x, s.box_matrix if triclinic else s.box, tm, qm, s.nbfp, s.excl_off, s.excl_idx), opt)
    h = fc.nb
    h.fep_set_atoms(tA, tB, qA, qB)
    oracle.set_triclinic(s.box_offdiag if triclinic else None)
    try:
        want = oracle.fep_pair_list(s.x, s.box, rlist, pert, s.excl_off, s.excl_idx)
        sv = oracle.shift_vectors(s.box).astype(np.float64)
    finally:
        oracle.set_triclinic(None)
    nri, nrj = h.fep_build_list()
    got = h.fep_list()
    assert nrj == len(want[3]) and got[2][-1] == nrj and nri == len(got[0])
    assert np.array_equal(_canonical(got), _canonical(want))
    assert len(np.unique(got[0].astype(np.int64) * 64 + got[1])) == nri and (np.diff(got[2]) > 0).all()  # one entry per (i, shift)
    assert set(np.unique(got[0])) <= set(np.nonzero(pert)[0])  # i-atoms are perturbed atoms
    res = []
    for lst in (None, want):
        if lst is not None:
            h.fep_upload_list(*lst)
        h.set_x(s.x)
        h.clear_outputs()
        h.fep_launch(**S.FEP_CASES["sc1coul"])
        f = h.get_f().astype(np.float64)
        fs = h.get_outputs()[0].astype(np.float64)
        res.append((f, np.einsum("sa,sb->ab", sv, fs), np.array(h.fep_outputs())))
    (f0, v0, o0), (f1, v1, o1) = res
    assert relrms(f0, f1) < 2e-6
    assert np.abs(v0 - v1).max() <= 2e-5 * np.abs(v1).max()
    assert np.abs(o0 - o1).max() <= 2e-5 * np.abs(o1).max()
    # as part of the step: compute() = cluster-pair kernel + free-energy kernel in one captured graph
    f_plain = fc.compute(s.x).astype(np.float64)
    h.fep_build_list()
    h.fep_outputs()
    h.fep_in_step(**S.FEP_CASES["sc1coul"])
    for rep in range(2):
        assert relrms(fc.compute(s.x).astype(np.float64) - f_plain, f0) < 1e-5 * max(1.0, np.abs(f_plain).max() / np.abs(f0).max())
    assert np.abs(np.array(h.fep_outputs()) - 2 * o0).max() <= 2e-5 * 2 * np.abs(o0).max()
    h.fep_in_step(None)
    assert relrms(fc.compute(s.x).astype(np.float64), f_plain) < 1e-6
    # twice the same list: reproducible order
    h.fep_build_list()
    again = h.fep_list()
    assert all(np.array_equal(a, b) for a, b in zip(got, again))
    fc.nb.close()


@pytest.mark.parametrize("case", ["sc1", "nosc", "sc2coul"])
def test_fep_kernel_potential_switch(built, case):
    """LJ potential switch 0.75 -> 0.9 nm in the free-energy kernel (applied on the soft-cored distance, nb_free_energy.cpp:613-625)
    against the oracle and the committed output of the reference kernel; reaction field."""
    S = g.systems
    s, pert, tA, tB, qA, qB, tm, qm = S.perturbed_water()
    kw = S.FEP_CASES[case]
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.ReactionField, computeVirialAndEnergy=True,
                            vdwModifier=g.VdwModifier.PotentialSwitch, vdwSwitch=0.75)
    fc = g.ForceCalculator(g.SimulationState(s.x, s.box, tm, qm, s.nbfp, s.excl_off, s.excl_idx), opt)
    h = fc.nb
    lst = oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx)
    h.fep_set_atoms(tA, tB, qA, qB)
    h.fep_upload_list(*lst)
    h.set_x(s.x)
    h.clear_outputs()
    h.fep_launch(**kw)
    f = h.get_f().astype(np.float64)
    out4 = np.array(h.fep_outputs())
    k, c = S.rf_constants(RC, eps_rf=1.0)
    fo, fso, o4 = oracle.fep_kernel(s.x, oracle.shift_vectors(s.box), s.nbfp, tA, tB, qA, qB, *lst, RC, k_rf=k, c_rf=c, rvdw_switch=0.75, **kw)
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_water_3k_fep_rf.npz"))
    assert relrms(f, fo.astype(np.float64)) < 1e-5 and relrms(f, gd["f_pswitch_" + case].astype(np.float64)) < 1e-5
    assert relrms(f, gd["f_" + case].astype(np.float64)) > 2e-5  # not the unswitched forces
    o4r = gd["out4_pswitch_" + case]
    assert np.abs(out4 - np.array(o4)).max() <= 2e-5 * np.abs(np.array(o4)).max()
    assert np.abs(out4 - o4r).max() <= 2e-5 * np.abs(o4r).max()
    fc.nb.close()


@pytest.mark.parametrize("rule,ljpme", [("geom", g.LjPme.Geometric), ("lb", g.LjPme.LorentzBerthelot)])
def test_fep_kernel_ljpme(built, rule, ljpme):
    """LJ-PME in the free-energy kernel (nb_free_energy.cpp:586-611 and :725-770: cut-off on the plain distance, the grid potential
    at the cut-off, the grid part of the dispersion taken off unsoftened -- also for excluded pairs and an atom with itself) with
    Ewald electrostatics, both grid combination rules, against the oracle and the committed outputs of the reference kernel
    (tests/golden/ref_water_3k_fep_ljpme.npz).  Oxygens and hydrogens carry LJ and disappear in state B."""
    S = g.systems
    s, pert, tA, tB, qA, qB, tm, qm, nbfp = S.perturbed_water_ljpme()
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_water_3k_fep_ljpme.npz"))
    beta, sh, bl, shlj = (float(gd[k]) for k in ("beta", "sh_ewald", "ewaldcoeff_lj", "sh_lj_ewald"))
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True, ewaldPotentialShift=True,
                            ljPme=ljpme, ljPmeEwaldCoeff=bl)
    fc = g.ForceCalculator(g.SimulationState(s.x, s.box, tm, qm, nbfp, s.excl_off, s.excl_idx), opt)
    h = fc.nb
    lst = oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx)
    h.fep_set_atoms(tA, tB, qA, qB)
    h.fep_upload_list(*lst)
    m = np.ones(45, bool)
    m[nb.CENTRAL] = False
    for case in ("sc1", "nosc", "sc2coul", "sc1coul"):
        kw = S.FEP_CASES[case]
        h.set_x(s.x)
        h.clear_outputs()
        h.fep_launch(**kw)
        f = h.get_f().astype(np.float64)
        fs = h.get_outputs()[0].astype(np.float64)
        out4 = np.array(h.fep_outputs())
        fo, fso, o4 = oracle.fep_kernel(s.x, oracle.shift_vectors(s.box), nbfp, tA, tB, qA, qB, *lst, RC, ewaldcoeff=beta, sh_ewald=sh,
                                        ljpme=ljpme.value, ewaldcoeff_lj=bl, sh_lj_ewald=shlj, **kw)
        tag = "_%s_%s" % (rule, case)
        assert relrms(f, fo.astype(np.float64)) < 1e-5 and relrms(f, gd["f" + tag].astype(np.float64)) < 1e-5
        fsr = gd["fshift" + tag].astype(np.float64)
        assert np.abs(fs[m] - fsr[m]).max() <= 2e-5 * np.abs(fsr[m]).max()
        o4r = gd["out4" + tag]
        assert np.abs(out4 - np.array(o4)).max() <= 2e-5 * np.abs(np.array(o4)).max()
        assert np.abs(out4 - o4r).max() <= 2e-5 * np.abs(o4r).max()
        if case == "sc1":
            assert relrms(f, gd["f_cut_sc1"].astype(np.float64)) > 1e-4  # not the cut-off LJ forces
    fc.nb.close()


def test_fep_kernel_twin_range(built):
    """rvdw = 0.8 < rcoulomb = 0.9 in the free-energy kernel (nb_free_energy.cpp:300-301, :564-587: the list is cut at the larger
    radius, each interaction at its own) for cut-off LJ, the potential switch 0.7 -> 0.8 and LJ-PME, Ewald electrostatics, against
    the oracle and the committed outputs of the reference kernel (tests/golden/ref_water_3k_fep_twin.npz)."""
    S = g.systems
    s, pert, tA, tB, qA, qB, tm, qm, nbfp = S.perturbed_water_ljpme()
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_water_3k_fep_twin.npz"))
    beta, sh, bl, shlj, rvdw = (float(gd[k]) for k in ("beta", "sh_ewald", "ewaldcoeff_lj", "sh_lj_ewald", "rvdw"))
    lst = oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx)
    for tag, rule, sw, case in S.FEP_TWIN:
        kw = S.FEP_CASES[case]
        opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True, ewaldPotentialShift=True,
                                vdwCutoff=rvdw, vdwSwitch=sw, vdwModifier=g.VdwModifier.PotentialSwitch if sw else g.VdwModifier.PotentialShift,
                                ljPme=g.LjPme(rule), ljPmeEwaldCoeff=bl if rule else 0.0)
        fc = g.ForceCalculator(g.SimulationState(s.x, s.box, tm, qm, nbfp, s.excl_off, s.excl_idx), opt)
        h = fc.nb
        h.fep_set_atoms(tA, tB, qA, qB)
        h.fep_upload_list(*lst)
        h.set_x(s.x)
        h.clear_outputs()
        h.fep_launch(**kw)
        f = h.get_f().astype(np.float64)
        out4 = np.array(h.fep_outputs())
        okw = dict(kw, ewaldcoeff=beta, sh_ewald=sh, rvdw=rvdw, rvdw_switch=sw)
        if rule:
            okw.update(ljpme=rule, ewaldcoeff_lj=bl, sh_lj_ewald=shlj)
        fo, fso, o4 = oracle.fep_kernel(s.x, oracle.shift_vectors(s.box), nbfp, tA, tB, qA, qB, *lst, RC, **okw)
        assert relrms(f, fo.astype(np.float64)) < 1e-5 and relrms(f, gd["f_" + tag].astype(np.float64)) < 1e-5, tag
        o4r = gd["out4_" + tag]
        assert np.abs(out4 - np.array(o4)).max() <= 2e-5 * np.abs(np.array(o4)).max(), tag
        assert np.abs(out4 - o4r).max() <= 2e-5 * np.abs(o4r).max(), tag
        fc.nb.close()


def test_fep_refuses_what_is_not_built(built):
    S = g.systems
    s, pert, tA, tB, qA, qB, tm, qm = S.perturbed_water()
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, vdwModifier=g.VdwModifier.ForceSwitch, vdwSwitch=0.75)
    fc = g.ForceCalculator(g.SimulationState(s.x, s.box, tm, qm, s.nbfp, s.excl_off, s.excl_idx), opt)
    fc.nb.fep_set_atoms(tA, tB, qA, qB)
    fc.nb.fep_upload_list(*oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx))
    with pytest.raises(nb.B200NBError):
        fc.nb.fep_launch(0.5, 0.5)  # LJ force switch: not in the reference's free-energy kernel, not built here
    with pytest.raises(nb.B200NBError):
        fc.nb.fep_upload_list([0], [22], [0, 1], [s.n + 5], [1])  # j-atom out of range
    with pytest.raises(nb.B200NBError):
        fc.nb.fep_launch(0.5, 0.5, sc_power=3)
    fc.nb.close()
