"""GPU parity of the domain-decomposed path on ONE GPU: the ranks are threads of this process, each with its own
NbnxmGpu context, exchanging halos through the in-process loopback transport (the multi-GPU run uses the same
DomainRank code over NCCL).  Global forces / pair set / energies / virial must equal the single-domain oracle."""
import os
import threading

import numpy as np
import pytest

import gmxapi_b200 as g
from gmxapi_b200 import lib as nb
from gmxapi_b200.domdec import DomainRank, LoopbackTransport
from oracle import oracle

pytestmark = pytest.mark.gpu
RC = 0.9
ENERGY_TOL = 2e-5  # relative, against the oracle's double-precision sums (the reference's own FP32 sums sit 1e-5 .. 2e-4 from those)
CENTRAL = 22


def shift_index(t):
    return 5 * (3 * (t[2] + 1) + (t[1] + 1)) + t[0] + 2


def canonical(pairs_global, tx_dd):
    """(i, j, shift) with the DD x-shift folded in, then brought to the half-list convention (shift <= CENTRAL,
    central pairs i < j) so that sets from different decompositions compare equal."""
    p = np.asarray(pairs_global, np.int64).reshape(-1, 3)
    s = p[:, 2]
    t = np.stack([(s % 5) - 2 + tx_dd, (s // 5) % 3 - 1, s // 15 - 1], 1)
    idx = 5 * (3 * (t[:, 2] + 1) + (t[:, 1] + 1)) + t[:, 0] + 2
    swap = (idx > CENTRAL) | ((idx == CENTRAL) & (p[:, 0] > p[:, 1]))
    i = np.where(swap, p[:, 1], p[:, 0])
    j = np.where(swap, p[:, 0], p[:, 1])
    t = np.where(swap[:, None], -t, t)
    idx = 5 * (3 * (t[:, 2] + 1) + (t[:, 1] + 1)) + t[:, 0] + 2
    return (i << 34) | (j << 6) | idx


def run_ranks(s, opt, nranks, flags, use_windows=True, nsteps=1, moved_x=None, repart_on_device=True):
    hub = LoopbackTransport(nranks)
    out = [None] * nranks
    err = []

    def work(r):
        try:
            d = DomainRank(s, opt, hub.endpoint(r), rank=r, nranks=nranks, device=0, use_windows=use_windows)
            assert d.use_windows == use_windows
            import torch
            xh = np.ascontiguousarray(s.x[d.plan.home])
            if r % 2 == 0:
                xh = torch.from_numpy(xh).pin_memory()  # even ranks: pinned host buffers, read / written in place by the kernels
            for _ in range(nsteps):  # repeated steps reuse the windows: flags advance, buffers are overwritten
                f, fs, elj, eel = d.compute(xh, flags)
            if moved_x is not None:
                # the atoms moved (integration is the caller's business): new coordinates of the atoms this rank owns go to
                # the device, then a repartitioning search step, then a step on the new decomposition
                old_home = d.plan.home.copy()
                plan = d.repartition(x_home=moved_x[old_home], on_device=repart_on_device)
                assert len(np.setdiff1d(plan.home, old_home)) > 0, "nothing migrated: the test does not test"
                from gmxapi_b200.domdec import wrap_into_box
                xh = np.ascontiguousarray(wrap_into_box(moved_x, s.box)[plan.home])
                assert np.array_equal(d.x[:plan.nhome].cpu().numpy(), xh)
                for _ in range(2):
                    f, fs, elj, eel = d.compute(xh, flags)
            pr = d.nb.pairs(RC)
            p = d.plan
            loc = p.local
            halo_periodic = p.recv_from_periodic
            is_halo = pr[:, 1] >= p.nhome
            gl = np.stack([loc[pr[:, 0]], loc[pr[:, 1]], pr[:, 2]], 1)
            keys = np.concatenate([canonical(gl[~is_halo], 0), canonical(gl[is_halo], -1 if halo_periodic else 0)])
            out[r] = dict(home=p.home, f=f.numpy().copy(), fs=fs, elj=elj, eel=eel, keys=keys, nhalo=p.nhalo, halo=p.halo.copy(),
                          send_local=p.send_local.copy(), npairs=len(pr))
            hub.endpoint(r).barrier()
            d.close()
        except Exception as e:  # noqa: BLE001
            err.append((r, repr(e)))
            try:
                hub._bar.abort()
            except Exception:
                pass

    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not err, err
    return out


@pytest.mark.parametrize("name,nranks,coulomb,windows", [("water_24k", 2, g.CoulombType.Pme, True),
                                                         ("water_24k", 3, g.CoulombType.ReactionField, True),
                                                         ("water_24k", 6, g.CoulombType.Pme, True),
                                                         ("water_96k", 4, g.CoulombType.Pme, True),
                                                         ("water_24k", 3, g.CoulombType.Pme, False)])
def test_domain_decomposition_matches_single_domain(built, name, nranks, coulomb, windows):
    """windows=True: halos move through the peer-memory windows of b200nb_dd_step (the product path);
    windows=False: through the transport's send/recv with the separate pack / unpack kernels."""
    s = g.systems.named(name)
    # reaction field with epsilon_rf = infinity (benchmark/bench_setup.cpp:152-155): the force vanishes at the cut-off, so
    # a pair flipped by the rounding of the periodic-edge shift (below) cannot show up in the forces
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=coulomb, computeVirialAndEnergy=True, epsilonRf=0.0)
    flags = nb.FLAG_ENERGY | nb.FLAG_VIRIAL
    res = run_ranks(s, opt, nranks, flags, use_windows=windows, nsteps=3 if windows else 1)
    if coulomb == g.CoulombType.Pme:
        kw = dict(eeltype=oracle.EEL_EWALD, beta=float(np.float32(g.systems.ewald_beta(RC))))
    else:
        k, c = g.systems.rf_constants(RC, eps_rf=0.0)
        kw = dict(eeltype=oracle.EEL_RF, k_rf=k, c_rf=c)
    fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, **kw)
    # pair set: every pair exactly once over all ranks, identical to the single-domain set
    keys = np.sort(np.concatenate([r["keys"] for r in res]))
    ok = oracle.canonical_pairs(oracle.pair_set(s.x, s.box, RC, s.excl_off, s.excl_idx))
    # Interior halos carry unshifted coordinates, so their pairs are bit-exact. On the periodic edge dd_move_x sends
    # x_j + box (rounded to float32, domdec/domdec.cpp:300-318) where the single-domain kernel evaluates
    # (x_i - box) - x_j, so a pair whose r^2 sits within rounding of rc^2 may flip -- the reference's own DD runs
    # differ from its single-rank runs in exactly this way. Any difference must be such a pair.
    diff = np.setxor1d(keys, ok)
    assert len(diff) <= max(2, npairs // 500000)
    sv = oracle.shift_vectors(s.box).astype(np.float64)
    for k in diff:
        i, j, sh = int(k >> 34), int((k >> 6) & ((1 << 28) - 1)), int(k & 63)
        assert (sh % 5) - 2 != 0, "a pair that does not cross the periodic x edge differs"
        r2 = ((s.x[i].astype(np.float64) + sv[sh] - s.x[j].astype(np.float64)) ** 2).sum()
        assert abs(r2 - RC * RC) < 5e-6
    assert abs(len(keys) - npairs) <= len(diff)
    # forces
    f = np.zeros((s.n, 3), np.float64)
    for r in res:
        assert r["nhalo"] > 0
        f[r["home"]] = r["f"]
    assert np.sqrt(((f - fo) ** 2).sum() / (fo ** 2).sum()) < 1e-5
    # energies
    elj, eel = sum(r["elj"] for r in res), sum(r["eel"] for r in res)
    assert abs(elj - evo) <= ENERGY_TOL * abs(evo)
    assert abs(eel - eco) <= ENERGY_TOL * abs(eco)
    # virial: -1/2 [ sum_a x_a (x) f_a + sum_s shift_vec[s] (x) fshift[s] ], decomposition-invariant
    fs = sum(r["fs"].astype(np.float64) for r in res)
    x = s.x.astype(np.float64)
    vir_g = -0.5 * (x.T @ f + sv.T @ fs)
    vir_o = -0.5 * (x.T @ fo + sv.T @ fso)
    assert np.abs(vir_g - vir_o).max() <= 1e-5 * np.abs(vir_o).max()


@pytest.mark.parametrize("nranks,windows", [(3, True), (2, True), (3, False)])
def test_repartition_after_motion(built, nranks, windows):
    """DD repartitioning (dd_partition_system): after the atoms moved -- here by a rigid translation across slab faces and
    the box edge plus a little noise -- DomainRank.repartition() migrates them, rebuilds halo plan, grids and lists, and the
    decomposed forces / pair set again equal the single-domain oracle on the new coordinates."""
    from gmxapi_b200.domdec import wrap_into_box
    s = g.systems.named("water_24k")
    rng = np.random.Generator(np.random.PCG64(23))
    x1 = (s.x + np.array([0.43, 0.31, -0.27], np.float32) + rng.uniform(-0.01, 0.01, s.x.shape)).astype(np.float32)
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True)
    flags = nb.FLAG_ENERGY | nb.FLAG_VIRIAL
    res = run_ranks(s, opt, nranks, flags, use_windows=windows, nsteps=1, moved_x=x1)
    x1w = wrap_into_box(x1, s.box)
    fo, fso, evo, eco, npairs = oracle.forces(x1w, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx,
                                              eeltype=oracle.EEL_EWALD, beta=float(np.float32(g.systems.ewald_beta(RC))))
    owned = np.concatenate([r["home"] for r in res])
    assert np.array_equal(np.sort(owned), np.arange(s.n))
    keys = np.sort(np.concatenate([r["keys"] for r in res]))
    ok = oracle.canonical_pairs(oracle.pair_set(x1w, s.box, RC, s.excl_off, s.excl_idx))
    diff = np.setxor1d(keys, ok)  # only pairs within rounding of rc^2 across the periodic x edge may differ (see above)
    assert len(diff) <= 2
    f = np.zeros((s.n, 3), np.float64)
    for r in res:
        f[r["home"]] = r["f"]
    assert np.sqrt(((f - fo) ** 2).sum() / (fo ** 2).sum()) < 1e-5
    elj, eel = sum(r["elj"] for r in res), sum(r["eel"] for r in res)
    assert abs(elj - evo) <= ENERGY_TOL * abs(evo)
    assert abs(eel - eco) <= ENERGY_TOL * abs(eco)


@pytest.mark.parametrize("nranks", [2, 3])
def test_device_repartition_equals_host_repartition(built, nranks):
    """The repartitioning kernels (csrc/dd_partition.cu: wrap + ownership, stable compaction, message packing, merge into the new
    home set, halo selection, local topology on the device) against the numpy restatement of the same step (migrate_atoms +
    DomainPlan.local_topology): identical home sets, send lists and halos on every rank, identical pair counts (the exclusions
    went through the device-side global -> local renumbering) and forces equal up to summation order."""
    s = g.systems.named("water_24k")
    rng = np.random.Generator(np.random.PCG64(29))
    x1 = (s.x + np.array([-0.52, 0.2, 0.35], np.float32) + rng.uniform(-0.01, 0.01, s.x.shape)).astype(np.float32)
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True)
    flags = nb.FLAG_ENERGY | nb.FLAG_VIRIAL
    dev = run_ranks(s, opt, nranks, flags, moved_x=x1, repart_on_device=True)
    host = run_ranks(s, opt, nranks, flags, moved_x=x1, repart_on_device=False)
    for a, b in zip(dev, host):
        assert np.array_equal(a["home"], b["home"])
        assert np.array_equal(a["send_local"], b["send_local"])
        assert np.array_equal(a["halo"], b["halo"])
        assert a["npairs"] == b["npairs"]
        assert np.array_equal(np.sort(a["keys"]), np.sort(b["keys"]))
        assert np.abs(a["f"] - b["f"]).max() <= 1e-4 * np.abs(b["f"]).max()
        assert abs(a["eel"] - b["eel"]) <= 1e-6 * abs(b["eel"])


# ---- 2-D / 3-D decomposition, half-shell rule (gmxapi_b200/domdec_nd.py) ----------------------------------------------------
_ND_CACHE = {}


def run_ranks_nd(grid, use_windows=True):
    """all ranks of a grid as threads on this GPU (loopback transport); cached: two tests look at the same run.
    use_windows: per-step halos through the peer-memory windows (b200nb_dd_set_links, the product path) or the transport"""
    key = (grid, use_windows)
    if key in _ND_CACHE:
        return _ND_CACHE[key]
    from gmxapi_b200.domdec_nd import DomainRankND
    s = g.systems.named("water_24k")
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True)
    flags = nb.FLAG_ENERGY | nb.FLAG_VIRIAL
    nranks = int(np.prod(grid))
    hub = LoopbackTransport(nranks)
    out = [None] * nranks
    err = []

    def work(r):
        try:
            d = DomainRankND(s, opt, hub.endpoint(r), grid, rank=r, device=0, use_windows=use_windows)
            assert d.use_windows == use_windows
            p = d.plan
            xh = np.ascontiguousarray(s.x[p.home])
            if use_windows and r % 2 == 0:
                import torch
                xh = torch.from_numpy(xh).pin_memory()  # even ranks: the kernels read / write pinned host buffers in place
            for _ in range(3 if use_windows else 2):  # repeated steps reuse the windows: flags advance
                f, fs, elj, eel = d.compute(xh, flags)
            pr = d.nb.pairs(RC)
            loc = p.local
            # shift (in box vectors) each halo atom was sent with; home atoms: none
            ksh = np.zeros((len(loc), 3), np.int64)
            off = p.nhome
            for rv in p.recv:
                ksh[off:off + len(rv["ids"])] = rv["shift"]
                off += len(rv["ids"])
            out[r] = dict(home=p.home, f=f.numpy().copy(), fs=fs, elj=elj, eel=eel, nhalo=p.nhalo,
                          pairs=np.stack([loc[pr[:, 0]], loc[pr[:, 1]], pr[:, 2]], 1), jshift=ksh[pr[:, 1]])
            hub.endpoint(r).barrier()
            d.close()
        except Exception as e:  # noqa: BLE001
            import traceback
            err.append((r, repr(e), traceback.format_exc()))
            try:
                hub._bar.abort()
            except Exception:
                pass

    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not err, err
    _ND_CACHE[key] = (s, out)
    return s, out


def canonical_nd(pairs, jshift):
    """(i, j, kernel shift) + the box-vector shift the j-atom was imported with -> half-list convention keys"""
    p = np.asarray(pairs, np.int64).reshape(-1, 3)
    sidx = p[:, 2]
    t = np.stack([(sidx % 5) - 2, (sidx // 5) % 3 - 1, sidx // 15 - 1], 1) - np.asarray(jshift, np.int64)  # shift of i relative to j
    idx = 5 * (3 * (t[:, 2] + 1) + (t[:, 1] + 1)) + t[:, 0] + 2
    swap = (idx > CENTRAL) | ((idx == CENTRAL) & (p[:, 0] > p[:, 1]))
    i = np.where(swap, p[:, 1], p[:, 0])
    j = np.where(swap, p[:, 0], p[:, 1])
    t = np.where(swap[:, None], -t, t)
    idx = 5 * (3 * (t[:, 2] + 1) + (t[:, 1] + 1)) + t[:, 0] + 2
    return (i << 34) | (j << 6) | idx


# (grid, per-step halos through the peer-memory windows?): the window path is the product path, the transport path stays covered
ND_GRIDS = [((2, 2, 1), True), ((2, 2, 2), True), ((3, 2, 1), True), ((2, 2, 1), False)]


@pytest.mark.parametrize("grid,windows", ND_GRIDS)
def test_decomposition_nd_matches_single_domain(built, grid, windows):
    """2 x 2, 2 x 2 x 2 and 3 x 2 ranks (half-shell rule, direct exchanges with every neighbour): pair set, forces and
    energies of the decomposed calculation equal the single-domain oracle."""
    s, res = run_ranks_nd(grid, windows)
    fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD,
                                              beta=float(np.float32(g.systems.ewald_beta(RC))))
    assert np.array_equal(np.sort(np.concatenate([r["home"] for r in res])), np.arange(s.n))
    keys = np.sort(np.concatenate([canonical_nd(r["pairs"], r["jshift"]) for r in res]))
    ok = oracle.canonical_pairs(oracle.pair_set(s.x, s.box, RC, s.excl_off, s.excl_idx))
    # images across a periodic edge arrive as x_j + box rounded to float32 (as dd_move_x sends them), where the single domain
    # evaluates (x_i - box) - x_j: only a pair within rounding of rc^2 may flip (see the slab test above)
    diff = np.setxor1d(keys, ok)
    assert len(diff) <= 6
    sv = oracle.shift_vectors(s.box).astype(np.float64)
    for k in diff:
        i, j, sh = int(k >> 34), int((k >> 6) & ((1 << 28) - 1)), int(k & 63)
        assert sh != CENTRAL
        r2 = ((s.x[i].astype(np.float64) + sv[sh] - s.x[j].astype(np.float64)) ** 2).sum()
        assert abs(r2 - RC * RC) < 5e-6
    assert len(keys) == len(np.unique(keys))
    f = np.zeros((s.n, 3), np.float64)
    for r in res:
        assert r["nhalo"] > 0
        f[r["home"]] = r["f"]
    assert np.sqrt(((f - fo) ** 2).sum() / (fo ** 2).sum()) < 1e-5
    elj, eel = sum(r["elj"] for r in res), sum(r["eel"] for r in res)
    assert abs(elj - evo) <= ENERGY_TOL * abs(evo)
    assert abs(eel - eco) <= ENERGY_TOL * abs(eco)


@pytest.mark.parametrize("grid,windows", ND_GRIDS)
def test_decomposition_nd_virial(built, grid, windows):
    """the virial -1/2 [ sum_a x_a (x) f_a + sum_s shift_vec[s] (x) fshift[s] ] is decomposition-invariant: forces on images
    that crossed a periodic edge enter the shift forces of the shift they were sent with (domdec.cpp:426-458)"""
    s, res = run_ranks_nd(grid, windows)
    fo, fso, _, _, _ = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD,
                                     beta=float(np.float32(g.systems.ewald_beta(RC))))
    sv = oracle.shift_vectors(s.box).astype(np.float64)
    f = np.zeros((s.n, 3), np.float64)
    for r in res:
        f[r["home"]] = r["f"]
    fs = sum(r["fs"].astype(np.float64) for r in res)
    x = s.x.astype(np.float64)
    vir_g = -0.5 * (x.T @ f + sv.T @ fs)
    vir_o = -0.5 * (x.T @ fo + sv.T @ fso)
    assert np.abs(vir_g - vir_o).max() <= 1e-5 * np.abs(vir_o).max()


@pytest.mark.parametrize("grid", [(2, 2, 1), (3, 1, 2)])
def test_repartition_nd_after_motion(built, grid):
    """repartitioning of a 2 x 2 and a 3 x 1 x 2 decomposition after a rigid translation across faces, an edge and the box boundary: on the
    device (csrc/dd_partition.cu: 28-way partition by neighbour offset, per-offset boundary selection, device-side local topology)
    and through the numpy restatement of the same step -- identical plans, forces against the single-domain oracle"""
    from gmxapi_b200.domdec import wrap_into_box
    from gmxapi_b200.domdec_nd import DomainRankND
    s = g.systems.named("water_24k")
    rng = np.random.Generator(np.random.PCG64(29))
    x1 = (s.x + np.array([0.43, -0.37, 0.21], np.float32) + rng.uniform(-0.01, 0.01, s.x.shape)).astype(np.float32)
    x1w = wrap_into_box(x1, s.box)
    opt = g.NBKernelOptions(pairlistCutoff=RC, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True)
    flags = nb.FLAG_ENERGY | nb.FLAG_VIRIAL
    nranks = int(np.prod(grid))

    def run(on_device):
        hub = LoopbackTransport(nranks)
        out, err = [None] * nranks, []

        def work(r):
            try:
                d = DomainRankND(s, opt, hub.endpoint(r), grid, rank=r, device=0)
                d.compute(np.ascontiguousarray(s.x[d.plan.home]), flags)
                old_home = d.plan.home.copy()
                plan = d.repartition(x_home=x1[old_home], on_device=on_device)
                assert len(np.setdiff1d(plan.home, old_home)) > 0
                assert np.array_equal(d.x[:plan.nhome].cpu().numpy(), x1w[plan.home])
                f, fs, elj, eel = d.compute(np.ascontiguousarray(x1w[plan.home]), flags)
                out[r] = dict(home=plan.home, f=f.numpy().copy(), elj=elj, eel=eel, halo=plan.halo.copy(),
                              send=[sd["local"].copy() for sd in plan.send], npairs=d.pair_count(RC))
                hub.endpoint(r).barrier()
                d.close()
            except Exception as e:  # noqa: BLE001
                import traceback
                err.append((r, repr(e), traceback.format_exc()))
                try:
                    hub._bar.abort()
                except Exception:
                    pass

        th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=300)
        assert not err, err
        return out

    out = run(True)
    host = run(False)
    for a, b in zip(out, host):
        assert np.array_equal(a["home"], b["home"]) and np.array_equal(a["halo"], b["halo"]) and a["npairs"] == b["npairs"]
        assert len(a["send"]) == len(b["send"]) and all(np.array_equal(u, v) for u, v in zip(a["send"], b["send"]))
        assert np.abs(a["f"] - b["f"]).max() <= 1e-4 * np.abs(b["f"]).max()
    fo, _, evo, eco, _ = oracle.forces(x1w, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD,
                                       beta=float(np.float32(g.systems.ewald_beta(RC))))
    assert np.array_equal(np.sort(np.concatenate([r["home"] for r in out])), np.arange(s.n))
    f = np.zeros((s.n, 3), np.float64)
    for r in out:
        f[r["home"]] = r["f"]
    assert np.sqrt(((f - fo) ** 2).sum() / (fo ** 2).sum()) < 1e-5
    assert abs(sum(r["elj"] for r in out) - evo) <= ENERGY_TOL * abs(evo)
    assert abs(sum(r["eel"] for r in out) - eco) <= ENERGY_TOL * abs(eco)
