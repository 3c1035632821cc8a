"""-m gpu: the listed ("bonded") interactions on the nonbonded buffers, gmxapi_b200/csrc/bonded.cu, through the C ABI, against the
restatement of the reference's GPU bonded kernels (oracle/nbnxm_oracle.c orc_bonded, pinned to the reference's CPU functions in
tests/test_oracle_cpu.py) and against the committed outputs of those functions (tests/golden/ref_bonded_chains.npz).
Tolerances: forces 1e-5 relative RMS (north_star), energies and shift forces 2e-5 of the largest."""
import os

import numpy as np
import pytest

import gmxapi_b200 as g
from gmxapi_b200 import lib as nb
from oracle import oracle

pytestmark = pytest.mark.gpu
RC = 0.9
SCALE14 = 138.935458 * 0.5


def relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def _calculator(s, box):
    n = s["n"]
    nbfp = np.array([[[6 * 2.0e-3, 12 * 2.0e-6]]], np.float32)
    # every atom excludes itself and its chain neighbours up to three bonds away (what a topology with 1-4 pairs generates)
    length = 24
    off, idx = [0], []
    for a in range(n):
        c0 = a - a % length
        idx += [b for b in range(max(c0, a - 3), min(c0 + length, a + 4))]
        off.append(len(idx))
    state = g.SimulationState(s["x"], box, np.zeros(n, np.int32), s["q"], nbfp, np.array(off, np.int32), np.array(idx, np.int32))
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.ReactionField, computeVirialAndEnergy=True)
    return g.ForceCalculator(state, opt)


@pytest.mark.parametrize("cell", ["rect", "tric"])
def test_bonded_kernel_matches_oracle_and_reference(built, cell):
    S = g.systems
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_bonded_chains.npz"))
    bm = None if cell == "rect" else gd["box_triclinic"]
    s = S.bonded_chains(box_matrix=bm)
    B = np.diag(s["box"]).astype(np.float32) if bm is None else np.asarray(bm, np.float32)
    fc = _calculator(s, s["box"] if bm is None else B)
    h = fc.nb
    flags = nb.FLAG_ENERGY | nb.FLAG_VIRIAL
    m = np.ones(45, bool)
    m[nb.CENTRAL] = False
    f_all = np.zeros((s["n"], 3))
    fs_all = np.zeros((45, 3))
    # one interaction type at a time: forces, shift forces, energy
    for kind in nb.BONDED_KINDS:
        for other in nb.BONDED_KINDS:
            h.bonded_set_list(other, np.zeros((0, nb.BONDED_NRAL[nb.BONDED_KINDS.index(other)] + 1), np.int32), np.zeros((0, 6), np.float32))
        d = s[kind]
        h.bonded_set_list(kind, d["iatoms"], d["params"])
        h.set_x(s["x"])
        h.clear_outputs()
        h.bonded_launch(flags, SCALE14)
        f = h.get_f().astype(np.float64)
        fs = h.get_outputs()[0].astype(np.float64)
        e = h.bonded_energies()
        fo, fso, eo = oracle.bonded(kind, d["iatoms"], d["params"], s["x"], s["q"], B, SCALE14)
        assert relrms(f, fo) < 1e-5, kind
        assert relrms(f, gd["f_%s_%s" % (cell, kind)].astype(np.float64)) < 1e-5, kind
        assert np.abs(fs[m] - fso[m]).max() <= 2e-5 * np.abs(fso).max(), kind
        assert abs(e[kind] - eo[0]) <= 2e-5 * abs(eo[0]), kind
        if kind == "lj14":
            assert abs(e["coul14"] - eo[1]) <= 2e-5 * abs(eo[1])
        else:
            assert abs(e[kind] - float(gd["e_%s_%s" % (cell, kind)])) <= 2e-5 * abs(eo[0]), kind
            fsr = gd["fshift_%s_%s" % (cell, kind)].astype(np.float64)
            assert np.abs(fs[m] - fsr[m]).max() <= 2e-5 * np.abs(fsr).max(), kind
        assert all(v == 0.0 for k, v in e.items() if k not in (kind, "coul14"))
        assert all(v == 0.0 for v in h.bonded_energies().values())  # read and reset
        f_all += fo
        fs_all += fso
    # all types in one launch, force-only flavour, on top of the nonbonded forces
    for kind in nb.BONDED_KINDS:
        h.bonded_set_list(kind, s[kind]["iatoms"], s[kind]["params"])
    h.set_x(s["x"])
    h.clear_outputs()
    h.launch_force(-1, 0)
    f_nb = h.get_f().astype(np.float64)
    h.set_x(s["x"])
    h.clear_outputs()
    h.launch_force(-1, 0)
    h.bonded_launch(0, SCALE14)
    f_sum = h.get_f().astype(np.float64)
    assert relrms(f_sum - f_nb, f_all) < 1e-5
    assert all(v == 0.0 for v in h.bonded_energies().values())  # no energies asked for
    h.set_x(s["x"])
    h.clear_outputs()
    h.bonded_launch(flags, SCALE14)
    fs = h.get_outputs()[0].astype(np.float64)
    assert relrms(h.get_f().astype(np.float64), f_all) < 1e-5
    assert np.abs(fs[m] - fs_all[m]).max() <= 2e-5 * np.abs(fs_all).max()
    # as part of the step: ForceCalculator.compute() = one captured graph with the bonded kernel between force kernel and un-sort
    f_plain = fc.compute(s["x"]).astype(np.float64)
    fs_plain = fc.shiftForces.astype(np.float64)
    assert relrms(f_plain, f_nb) < 1e-6
    h.bonded_energies()
    h.bonded_in_step(True, SCALE14)
    for rep in range(3):  # the replayed graph gives the same forces every time
        f_step = fc.compute(s["x"]).astype(np.float64)
        assert relrms(f_step - f_plain, f_all) < 1e-5, rep
        assert np.abs((fc.shiftForces.astype(np.float64) - fs_plain)[m] - fs_all[m]).max() <= 2e-5 * np.abs(fs_all).max()
    e = h.bonded_energies()  # summed over the three steps
    eo = oracle.bonded("bonds", s["bonds"]["iatoms"], s["bonds"]["params"], s["x"], s["q"], B, SCALE14)[2][0]
    assert abs(e["bonds"] - 3 * eo) <= 2e-5 * 3 * abs(eo)
    h.bonded_set_list("lj14", np.zeros((0, 3), np.int32), np.zeros((0, 6), np.float32))  # a changed list re-captures the step
    f_step = fc.compute(s["x"]).astype(np.float64)
    f14 = oracle.bonded("lj14", s["lj14"]["iatoms"], s["lj14"]["params"], s["x"], s["q"], B, SCALE14)[0]
    assert relrms(f_step - f_plain, f_all - f14) < 1e-5
    h.bonded_in_step(False)
    assert relrms(fc.compute(s["x"]).astype(np.float64), f_plain) < 1e-6
    fc.nb.close()


def test_bonded_input_checks(built):
    S = g.systems
    s = S.bonded_chains()
    fc = _calculator(s, s["box"])
    h = fc.nb
    with pytest.raises(nb.B200NBError):
        h.bonded_set_list("bonds", [[0, 0, s["n"]]], s["bonds"]["params"])  # atom out of range
    with pytest.raises(nb.B200NBError):
        h.bonded_set_list("angles", [[7, 0, 1, 2]], s["angles"]["params"])  # parameter index out of range
    with pytest.raises(nb.B200NBError):
        h._check(h._L.b200nb_bonded_set_list(h._h, 11, 0, None, 0, None), "bonded_set_list")
    h.bonded_launch(0, SCALE14)  # no lists: nothing to do
    assert all(v == 0.0 for v in h.bonded_energies().values())
    fc.nb.close()
