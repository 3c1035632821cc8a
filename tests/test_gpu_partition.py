"""-m gpu: the repartitioning kernels of gmxapi_b200/csrc/dd_partition.cu one by one, through the C ABI, against numpy on the same
inputs (bit-exact: these are integer / index kernels and float32 comparisons).  The decomposed runs that use them end to end are in
tests/test_gpu_domdec.py; this file is also what compute-sanitizer is pointed at (a decomposed step cannot run under the
sanitizer's kernel serialisation: its flag waits need the neighbour's kernels to run beside them)."""
import numpy as np
import pytest

import gmxapi_b200 as g
from gmxapi_b200 import lib as nb
from gmxapi_b200.domdec import DomainPlan, wrap_into_box
from gmxapi_b200.domdec_nd import DomainPlanND

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx(built):
    import torch
    h = nb.NbnxmGpu(0)
    yield h, torch
    h.synchronize()
    h.close()


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_wrap_classify_partition_pack_merge_slabs(ctx):
    h, torch = ctx
    rng = np.random.Generator(np.random.PCG64(5))
    box = np.array([6.2, 5.1, 4.3], np.float32)
    nranks, rank, n = 4, 1, 20011
    bounds = DomainPlan.boundaries(box, nranks)
    # atoms of slab 1 after a move: most stay, some cross either face, a few leave the box in y / z, two jump two slabs
    x = rng.uniform([bounds[1], 0, 0], [bounds[2], box[1], box[2]], (n, 3)).astype(np.float32)
    x[:, 0] += rng.normal(0, 0.12, n).astype(np.float32)
    x[::97, 1] += box[1]
    x[::89, 2] -= 2 * box[2]
    x[5, 0] = bounds[3] + 0.3
    x[6, 0] = bounds[0] + 0.1 + box[0]  # wraps to slab 0: the left neighbour
    x[7, 0] = bounds[3] + 0.01 + box[0] * 2  # wraps to slab 3: two away
    gid = np.sort(rng.choice(10 * n, n, replace=False)).astype(np.int32)
    xd, gd = _dev(torch, x), _dev(torch, gid)
    code = torch.empty(n, dtype=torch.int32, device="cuda")
    idx = torch.empty(n, dtype=torch.int32, device="cuda")
    h.dd_wrap_classify(xd.data_ptr(), n, box, bounds, nranks, rank, code.data_ptr())
    xw = wrap_into_box(x, box)
    assert np.array_equal(xd.cpu().numpy(), xw)
    owner = DomainPlan.owner_of(xw, box, nranks)
    want = np.where(owner == rank, 0, np.where(owner == rank - 1, 1, np.where(owner == rank + 1, 2, 3))).astype(np.int32)
    assert np.array_equal(code.cpu().numpy(), want) and (want == 3).sum() >= 2
    cnt = h.dd_partition_indices(code.data_ptr(), n, 4, idx.data_ptr())
    assert cnt == [int((want == k).sum()) for k in range(4)]
    assert np.array_equal(idx.cpu().numpy(), np.concatenate([np.nonzero(want == k)[0] for k in range(4)]))  # stable
    # leavers to the left as a message, then pretend they are what arrives (shuffled) at a rank that kept `stay`
    nstay, nl = cnt[0], cnt[1]
    out4 = torch.empty((nl, 4), dtype=torch.int32, device="cuda")
    h.dd_pack_atoms(idx.data_ptr() + 4 * nstay, nl, gd.data_ptr(), xd.data_ptr(), out4.data_ptr())
    m = out4.cpu().numpy()
    left = np.nonzero(want == 1)[0]
    assert np.array_equal(m[:, 0], gid[left]) and np.array_equal(m[:, 1:].view(np.float32), xw[left])
    perm = rng.permutation(nl)
    arrived = _dev(torch, m[perm])
    gout = torch.empty(nstay + nl, dtype=torch.int32, device="cuda")
    xout = torch.empty((nstay + nl, 3), dtype=torch.float32, device="cuda")
    h.dd_merge_home(idx.data_ptr(), nstay, gd.data_ptr(), xd.data_ptr(), arrived.data_ptr(), nl, gout.data_ptr(), xout.data_ptr())
    stay = np.nonzero(want == 0)[0]
    ids = np.concatenate([gid[stay], gid[left]])
    xs = np.concatenate([xw[stay], xw[left]])
    order = np.argsort(ids, kind="stable")
    assert np.array_equal(gout.cpu().numpy(), ids[order]) and np.array_equal(xout.cpu().numpy(), xs[order])
    # halo selection + gather
    lo, rlist = float(bounds[rank]), 0.9
    h.dd_select_lower_face(xout.data_ptr(), nstay + nl, lo, rlist, code.data_ptr())
    sel = DomainPlan._send_list(xs[order], np.arange(nstay + nl), bounds[rank], rlist)
    nkeep, nsend = h.dd_partition_indices(code.data_ptr(), nstay + nl, 2, idx.data_ptr())
    assert nsend == len(sel) and np.array_equal(idx[nkeep:nkeep + nsend].cpu().numpy(), sel)
    sg = torch.empty(nsend, dtype=torch.int32, device="cuda")
    h.dd_gather_int(idx.data_ptr() + 4 * nkeep, nsend, gout.data_ptr(), sg.data_ptr())
    assert np.array_equal(sg.cpu().numpy(), ids[order][sel])
    # empty inputs are legal
    assert h.dd_partition_indices(code.data_ptr(), 0, 4, idx.data_ptr()) == [0, 0, 0, 0]
    h.dd_merge_home(idx.data_ptr(), 0, gd.data_ptr(), xd.data_ptr(), 0, 0, gout.data_ptr(), xout.data_ptr())


@pytest.mark.parametrize("grid,rank", [((2, 2, 2), 5), ((3, 1, 2), 2), ((4, 3, 1), 7)])
def test_wrap_classify_and_boundary_selection_nd(ctx, grid, rank):
    h, torch = ctx
    rng = np.random.Generator(np.random.PCG64(11))
    box = np.array([6.2, 7.4, 5.6], np.float32)
    geo = DomainPlanND.__new__(DomainPlanND)
    geo._geometry(box, grid, rank, 0.9)
    n = 15013
    x = rng.uniform(geo.lo, geo.hi, (n, 3)).astype(np.float32)
    x += rng.normal(0, 0.15, (n, 3)).astype(np.float32)
    x[::101] += box * np.array([1, -1, 2], np.float32)
    xd = _dev(torch, x)
    code = torch.empty(n, dtype=torch.int32, device="cuda")
    idx = torch.empty(n, dtype=torch.int32, device="cuda")
    h.dd_wrap_classify_nd(xd.data_ptr(), n, box, grid, geo.coords, code.data_ptr())
    xw = wrap_into_box(x, box)
    assert np.array_equal(xd.cpu().numpy(), xw)
    owner = DomainPlanND.owner_of(xw, box, grid)
    got = code.cpu().numpy()
    for a in range(0, n, 7):
        c = int(got[a])
        oc = DomainPlanND.coords_of(int(owner[a]), grid)
        if c == 27:
            assert any(min((oc[d] - geo.coords[d]) % grid[d], (geo.coords[d] - oc[d]) % grid[d]) > 1 for d in range(3))
            continue
        o = (c // 9 - 1, (c // 3) % 3 - 1, c % 3 - 1)
        assert geo.neighbour(geo.coords, o, +1)[0] == owner[a], (a, c, o)
    assert (got == 13).sum() == (owner == rank).sum()
    cnt = h.dd_partition_indices(code.data_ptr(), n, 28, idx.data_ptr())
    assert cnt == [int((got == k).sum()) for k in range(28)]
    assert np.array_equal(idx.cpu().numpy(), np.concatenate([np.nonzero(got == k)[0] for k in range(28)]))
    for o in geo.offsets:
        h.dd_select_boundary(xd.data_ptr(), n, geo.lo, geo.hi, o, 0.9, code.data_ptr())
        want = np.zeros(n, np.int32)
        want[geo.boundary_atoms(xw, geo.coords, o)] = 1
        assert np.array_equal(code.cpu().numpy(), want), o


def test_local_topology_on_device_gives_the_same_pairs(ctx):
    """b200nb_dd_set_global_topology + b200nb_dd_set_local_atoms against b200nb_set_atoms with the numpy-built local topology: the
    same in-range, non-excluded pair set (the exclusions went through the device-side global -> local renumbering)."""
    h, torch = ctx
    s = g.systems.named("water_3k")
    opt = g.NBKernelOptions(pairlistCutoff=0.9, coulombType=g.CoulombType.Pme)
    from gmxapi_b200.nblib import configure_interactions
    # "local" atoms: a shuffled subset of the box (whole molecules and broken ones), as home + halo would be
    rng = np.random.Generator(np.random.PCG64(3))
    local = np.sort(rng.choice(s.n, 2400, replace=False)).astype(np.int32)
    plan = DomainPlan.from_parts(s.box, 1, 0, 0.9, local, np.zeros(0, np.int32), np.zeros(0, np.int32))
    types, q, eo, ei = plan.local_topology(s.types, s.q, s.excl_off, s.excl_idx)
    pairs = []
    for device_side in (False, True):
        configure_interactions(h, s.nbfp, opt, 0.9)
        h.set_box(s.box)
        if device_side:
            h.dd_set_global_topology(s.types, s.q, s.excl_off, s.excl_idx)
            lg = _dev(torch, local)
            h.dd_set_local_atoms(lg.data_ptr(), len(local))
        else:
            h.set_atoms(types, q, eo, ei)
        x = np.ascontiguousarray(s.x[local])
        h.put_on_grid(x, np.zeros(3, np.float32), np.asarray(s.box, np.float32), 0, 0, len(local))
        h.build_pairlist()
        p = h.pairs(0.9)
        pairs.append(np.sort((p[:, 0].astype(np.int64) << 34) | (p[:, 1].astype(np.int64) << 6) | p[:, 2]))
    assert len(pairs[0]) > 100000 and np.array_equal(pairs[0], pairs[1])
