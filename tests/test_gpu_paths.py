"""GPU tests of the paths round 1 left without one: b200nb_get_f(accumulate = 1), LJ switch
modifiers with reaction-field electrostatics, the copy-engine variant of b200nb_compute (B200NB_HOST_DMA=1), a rebuild without
the halo grid after one with it, and the reference's liquid benchmark water instead of the lattice."""
import os
import subprocess
import sys

import numpy as np
import pytest

import gmxapi_b200 as g
from gmxapi_b200 import lib as nb
from oracle import oracle

pytestmark = pytest.mark.gpu
RC = 0.9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relrms(a, b):
    return float(np.sqrt(((np.asarray(a, np.float64) - b) ** 2).sum() / (np.asarray(b, np.float64) ** 2).sum()))


def test_get_f_accumulate(built):
    """reduceKernel<accumulate = true> (mdlib/gpuforcereduction_impl.cu:70-104): f_out[a] += f_grid[cell[a]] on a device buffer"""
    import torch
    s = g.systems.named("water_3k")
    fc = g.ForceCalculator(g.SimulationState.from_system(s), g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme))
    f = fc.compute()
    h = fc.nb
    base = torch.from_numpy(np.random.default_rng(1).normal(size=(s.n, 3)).astype(np.float32)).cuda()
    acc = base.clone()
    h.get_f(acc.data_ptr(), on_device=True, accumulate=True)
    h.synchronize()
    assert np.allclose(acc.cpu().numpy(), base.cpu().numpy() + f, rtol=0, atol=1e-4 * np.abs(f).max())
    # a sub-range accumulates only there
    acc2 = base.clone()
    h.get_f(acc2.data_ptr(), on_device=True, accumulate=True, atom_begin=100, atom_end=200)
    h.synchronize()
    d = acc2.cpu().numpy() - base.cpu().numpy()
    assert np.abs(d[:100]).max() == 0 and np.abs(d[200:]).max() == 0 and np.allclose(d[100:200], f[100:200], atol=1e-4 * np.abs(f).max())
    with pytest.raises(nb.B200NBError):
        h.get_f(np.zeros((s.n, 3), np.float32), on_device=False, accumulate=True)  # host buffers are overwritten, never added to


@pytest.mark.parametrize("mod,rsw", [(g.VdwModifier.ForceSwitch, 0.75), (g.VdwModifier.PotentialSwitch, 0.75)])
def test_switch_modifier_with_reaction_field(built, mod, rsw):
    """the general kernels with RF electrostatics (the oracle is pinned to the live reference for this combination,
    tests/test_oracle_cpu.py)"""
    s = g.systems.named("water_3k")
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.ReactionField, computeVirialAndEnergy=True, epsilonRf=0.0,
                            vdwModifier=mod, vdwSwitch=rsw)
    fc = g.ForceCalculator(g.SimulationState.from_system(s), opt)
    f = fc.compute()
    k, c = g.systems.rf_constants(RC, eps_rf=0.0)
    fo, fso, evo, eco, _ = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_RF, k_rf=k, c_rf=c,
                                         vdw_modifier=mod.value, rvdw=RC, rvdw_switch=rsw)
    assert relrms(f, fo) < 1e-5
    elj, eel = fc.energies
    # switched LJ energy: near-cancelling sum, same absolute bar as the plain potential's 113 kJ/mol (see test_vdw_flavours)
    assert abs(elj - evo) <= 2e-5 * max(abs(evo), 113.0) and abs(eel - eco) <= 2e-5 * abs(eco)


def test_host_dma_variant_matches(built):
    """B200NB_HOST_DMA=1: b200nb_compute stages through the copy engines instead of letting the kernels read / write the pinned
    host buffers; chosen once per context from the environment, so the comparison runs in a child process"""
    code = ("import numpy as np, gmxapi_b200 as g\n"
            "s = g.systems.named('water_3k')\n"
            "fc = g.ForceCalculator(g.SimulationState.from_system(s), g.NBKernelOptions(pairlistCutoff=0.9, coulombType=g.CoulombType.Pme))\n"
            "np.save(r'%s', fc.compute())\n")
    out = {}
    for dma in ("0", "1"):
        path = os.path.join("/tmp", "b200nb_dma%s_%d.npy" % (dma, os.getpid()))
        r = subprocess.run([sys.executable, "-c", code % path], env=dict(os.environ, B200NB_HOST_DMA=dma, PYTHONPATH=ROOT),
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[dma] = np.load(path)
        os.remove(path)
    assert relrms(out["1"], out["0"]) < 1e-6  # same kernels, same list: only float-atomic ordering differs


def test_rebuild_without_halo_grid_drops_the_nonlocal_list(built):
    """a context that searched with a halo grid and is then re-gridded without one must not keep (and run) the old
    non-local list"""
    s = g.systems.named("water_3k")
    h = nb.NbnxmGpu(0)
    h.set_params(s.nbfp, RC, eeltype=nb.EEL_RF, k_rf=0.1, c_rf=1.0)
    n_home = 2400
    h.set_atoms(s.types, s.q, s.excl_off, s.excl_idx)
    h.set_box(s.box, pbc=(0, 1, 1))
    lo, hi = np.zeros(3, np.float32), np.asarray(s.box, np.float32)
    order = np.argsort(s.x[:, 0], kind="stable")
    # home = the 2400 atoms with the smallest x (whole molecules do not matter here), halo = the rest
    x = s.x[order]
    t, q = s.types[order], s.q[order]
    h.set_atoms(t, q)
    split = float(x[n_home, 0])
    h.put_on_grid(x, lo, [split, hi[1], hi[2]], 0, 0, n_home)
    h.put_on_grid(x, [split, 0, 0], hi, 1, n_home, s.n)
    h.build_pairlist()
    assert h.stats()["nentries_nonlocal"] > 0
    # same context, now a single domain of the first 2400 atoms only
    h.set_atoms(t[:n_home], q[:n_home])
    h.put_on_grid(x[:n_home], lo, [split, hi[1], hi[2]], 0, 0, n_home)
    h.build_pairlist()
    assert h.stats()["nentries_nonlocal"] == 0
    h.set_x(x[:n_home])
    h.clear_outputs()
    h.launch_force(-1, 0)  # would read stale slots through the old non-local list
    f = h.get_f()
    assert np.all(np.isfinite(f))
    h.close()


def test_tabulated_ewald(built):
    """the EL_EWALD_TAB flavour (b200nb_set_ewald_table): with the REFERENCE's own table against the reference's GPU-layout kernel,
    which interpolates the same table (kernel_gpu_ref.cpp:262-272), where oracle/_ref is on this box; with the analytic table of
    the Python mirror against the oracle's analytical Ewald (interpolation error of a 2000-point-per-nm table: < 2e-6)"""
    from oracle import gmxref
    s = g.systems.named("water_3k")
    beta = float(np.float32(g.systems.ewald_beta(RC)))
    fo, fso, evo, eco, _ = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD, beta=beta)
    fc = g.ForceCalculator(g.SimulationState.from_system(s), g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme,
                                                                                computeVirialAndEnergy=True))
    f_ana = fc.compute().copy()
    for energy in (True, False):
        ft = g.ForceCalculator(g.SimulationState.from_system(s), g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme,
                                                                                    computeVirialAndEnergy=energy, useTabulatedEwaldCorr=True))
        f_tab = ft.compute().copy()
        assert not np.array_equal(f_tab, f_ana)  # the other kernel ran
        assert relrms(f_tab, fo) < 1e-5
        if energy:
            assert abs(ft.energies[1] - eco) <= 2e-5 * abs(eco)  # energies use the analytical form in the tabulated kernels too
    if gmxref.available():
        r = gmxref.RefNbnxm(s.x, s.box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx, rc=RC, eeltype=gmxref.EEL_EWALD_TAB, ewaldcoeff=beta,
                            kernel=gmxref.KERNEL_GPUREF, nthreads=1)
        f_ref = r.compute(energy=False, virial=False)[0]
        table, scale = r.ewald_table()
        fc.nb.set_ewald_table(table, scale)
        f_tab = fc.compute()
        assert relrms(f_tab, f_ref) < 3e-6  # same table, same interpolation: closer than to the analytical form
        fc.nb.set_ewald_table(None, 0.0)
        assert relrms(fc.compute(), f_ana) < 1e-6  # and back
        r.close()


@pytest.mark.parametrize("name", ["ref_water_3k", "ref_water_24k"])
def test_reference_liquid_water(built, name):
    """BenchmarkSystem's own coordinates (nbnxm/benchmark/bench_coords.h through gmxapi_b200/data/ref_water_1000.npz): liquid
    structure instead of the jittered lattice -- grid order, pair set, forces, energies against the oracle"""
    s = g.systems.named(name)
    fc = g.ForceCalculator(g.SimulationState.from_system(s), g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme,
                                                                                computeVirialAndEnergy=True))
    f = fc.compute()
    go = oracle.put_on_grid(s.x, s.box)
    assert np.array_equal(fc.nb.grid_order(), go["atom_index"])
    beta = float(np.float32(g.systems.ewald_beta(RC)))
    fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD, beta=beta)
    gp = oracle.canonical_pairs(fc.nb.pairs(RC))
    op = oracle.canonical_pairs(oracle.pair_set(s.x, s.box, RC, s.excl_off, s.excl_idx))
    assert np.array_equal(gp, op)
    assert relrms(f, fo) < 1e-5
    elj, eel = fc.energies
    assert abs(elj - evo) <= 2e-5 * abs(evo) and abs(eel - eco) <= 2e-5 * abs(eco)


@pytest.mark.gpu
def test_describe_names_the_setup(built):
    """b200nb_describe: the one-line set-up summary (device, atoms, grid, flavour, list sizes) a caller logs."""
    import gmxapi_b200 as g
    s = g.systems.named("water_3k")
    fc = g.ForceCalculator(g.SimulationState.from_system(s), g.NBKernelOptions(pairlistCutoff=0.9, coulombType=g.CoulombType.Pme))
    line = fc.nb.describe()
    assert "3000 atoms" in line and "Ewald(analytical)" in line and "half-entries" in line and "sm_100" in line, line
    st = fc.nb.stats()
    assert "%d + 0 cluster pairs" % st["ntiles_outer"] in line, line
