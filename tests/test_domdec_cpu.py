"""CPU tests of the domain-decomposition host logic: the plan (who owns / sends / receives which atom) and
the halo exchange pattern over torch.distributed with the gloo backend, world_size 2 and 3 (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gmxapi_b200 import systems as S
from gmxapi_b200.domdec import DomainPlan, TorchDistTransport, migrate_atoms, wrap_into_box
from gmxapi_b200.domdec_nd import DomainPlanND, half_shell_offsets, migrate_atoms_nd

RLIST = 0.9


def _pairs_bruteforce(xi, xj, r, same):
    """index pairs (a, b) with |xi[a] - xj[b]| < r (no PBC: the caller has applied the shifts)."""
    out = []
    for a0 in range(0, len(xi), 512):
        d = xi[a0:a0 + 512, None, :].astype(np.float64) - xj[None, :, :].astype(np.float64)
        m = (d ** 2).sum(-1) < r * r
        a, b = np.nonzero(m)
        a = a + a0
        if same:
            k = a < b
            a, b = a[k], b[k]
        out.append(np.stack([a, b], 1))
    return np.concatenate(out) if out else np.zeros((0, 2), np.int64)


@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_plan_partitions_atoms_and_pairs(nranks):
    """Every atom has exactly one owner; the union over ranks of home-home and home-halo pairs within rlist along
    the decomposed dimension is every pair of the periodic system exactly once (eighth-shell rule,
    domdec/domdec.cpp:133-146)."""
    s = S.water_box(10, 6, 6, seed=3)  # 3.1 x 1.86 x 1.86 nm: x is the only dimension with > 2 rlist... y,z not periodic here
    x = s.x
    box = s.box
    plans = [DomainPlan(x, box, nranks, r, RLIST) for r in range(nranks)]
    owned = np.concatenate([p.home for p in plans])
    assert np.array_equal(np.sort(owned), np.arange(s.n))
    if nranks == 1:
        assert plans[0].nhalo == 0
        return
    # reference: all pairs within rlist with periodicity along x only (minimum image in x)
    d = x[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)
    d[..., 0] -= np.round(d[..., 0] / box[0]) * box[0]
    ref = np.argwhere(np.triu((d ** 2).sum(-1) < RLIST ** 2, 1))
    ref_keys = np.sort(ref[:, 0].astype(np.int64) * s.n + ref[:, 1])
    got = []
    for p in plans:
        xh = x[p.home]
        hh = _pairs_bruteforce(xh, xh, RLIST, True)
        got.append(np.stack([p.home[hh[:, 0]], p.home[hh[:, 1]]], 1))
        if p.nhalo:
            hx = p.halo_x(x)
            hj = _pairs_bruteforce(xh, hx, RLIST, False)
            got.append(np.stack([p.home[hj[:, 0]], p.halo[hj[:, 1]]], 1))
    got = np.concatenate(got)
    lo, hi = np.minimum(got[:, 0], got[:, 1]), np.maximum(got[:, 0], got[:, 1])
    keys = np.sort(lo.astype(np.int64) * s.n + hi)
    if nranks == 2:
        # with two ranks a pair can be in range both directly and through the periodic image only if box < 2 rlist
        assert box[0] >= 2 * RLIST
    assert len(keys) == len(np.unique(keys)), "a pair is computed on two ranks"
    assert np.array_equal(keys, ref_keys)


def test_plan_local_topology():
    s = S.named("water_3k")
    p = DomainPlan(s.x, s.box, 2, 0, RLIST)
    t, q, off, idx = p.local_topology(s.types, s.q, s.excl_off, s.excl_idx)
    n = p.nhome + p.nhalo
    assert len(t) == len(q) == n and len(off) == n + 1 and off[-1] == len(idx)
    assert np.array_equal(t, s.types[p.local]) and np.array_equal(q, s.q[p.local])
    # every local exclusion maps back to a global exclusion of the same atom, self included
    for a in (0, 1, p.nhome - 1, p.nhome, n - 1):
        ga = p.local[a]
        glob = set(s.excl_idx[s.excl_off[ga]:s.excl_off[ga + 1]].tolist())
        loc = set(p.local[idx[off[a]:off[a + 1]]].tolist())
        assert ga in loc and loc <= glob
        assert loc == {g for g in glob if g in set(p.local.tolist())}


def test_plan_rejects_thin_domains():
    s = S.named("water_3k")  # 3.1 nm box
    from gmxapi_b200.nblib import InputException
    with pytest.raises(InputException):
        DomainPlan(s.x, s.box, 4, 0, RLIST)  # 0.78 nm slabs < rlist


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = S.water_box(12, 6, 6, seed=5)
        plan = DomainPlan(s.x, s.box, world, rank, RLIST)
        t = TorchDistTransport()
        assert (t.rank, t.nranks) == (rank, world)
        # dd_move_x: pack on the sender (the CUDA pack kernel's arithmetic, restated with torch on the CPU for this
        # host-logic test only), transfer, compare with what the plan says the halo must hold
        x = torch.from_numpy(s.x[plan.local].copy())
        x[plan.nhome:] = float("nan")
        send = x[:plan.nhome][torch.from_numpy(plan.send_local).long()] + torch.from_numpy(plan.send_shift)
        t.sendrecv(send.contiguous(), plan.left, x[plan.nhome:], plan.right)
        ok_x = bool(np.array_equal(x[plan.nhome:].numpy(), plan.halo_x(s.x)))
        # dd_move_f: each halo atom carries f = its global index; after the return trip the owner must have received
        # exactly one contribution per sent atom, equal to that atom's global index
        f = torch.zeros((plan.nhome + plan.nhalo, 3))
        f[plan.nhome:] = torch.from_numpy(plan.halo.astype(np.float32))[:, None]
        recv = torch.zeros((len(plan.send_local), 3))
        t.sendrecv(f[plan.nhome:].contiguous(), plan.right, recv, plan.left)
        f[:plan.nhome].index_add_(0, torch.from_numpy(plan.send_local).long(), recv)
        exp = np.zeros(plan.nhome, np.float32)
        exp[plan.send_local] = plan.home[plan.send_local]
        ok_f = bool(np.array_equal(f[:plan.nhome, 0].numpy(), exp))
        tot = t.allreduce_sum(torch.tensor([float(plan.nhome)]))
        # window set-up exchange of DomainRank._open_windows: every rank learns both neighbours' (pid, handle, size);
        # what a rank will receive as halo must fit what its +x neighbour says it sends, and vice versa
        info = t.allgather_object(dict(pid=os.getpid(), rank=rank, nsend=len(plan.send_local), nhalo=plan.nhalo))
        ok_w = (info[plan.right]["nsend"] == plan.nhalo and info[plan.left]["nhalo"] == len(plan.send_local)
                and [i["rank"] for i in info] == list(range(world)) and len({i["pid"] for i in info}) == world)
        q.put((rank, ok_x, ok_f and ok_w, int(tot.item()) == s.n, plan.nhalo))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_x, ok_f, ok_n, nhalo in res:
        assert ok_x and ok_f and ok_n and nhalo > 0, (rank, ok_x, ok_f, ok_n, nhalo)


def _moved(s, seed=11, amp=0.35):
    """the same displaced coordinates on every rank: up to `amp` nm along each axis, so atoms cross slab faces and box edges"""
    rng = np.random.Generator(np.random.PCG64(seed))
    return (s.x + rng.uniform(-amp, amp, s.x.shape)).astype(np.float32)


def _migrate_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = S.water_box(12, 6, 6, seed=5)
        t = TorchDistTransport()
        plan0 = DomainPlan(s.x, s.box, world, rank, RLIST)
        x1 = _moved(s)
        home, xh, send_local, halo = migrate_atoms(t, s.box, world, rank, RLIST, plan0.home, x1[plan0.home])
        x1w = wrap_into_box(x1, s.box)
        exp = DomainPlan(x1w, s.box, world, rank, RLIST)  # what a plan made from the global coordinates says
        ok = (np.array_equal(home, exp.home) and np.array_equal(xh, x1w[exp.home]) and np.array_equal(send_local, exp.send_local)
              and np.array_equal(halo, exp.halo))
        plan1 = DomainPlan.from_parts(s.box, world, rank, RLIST, home, send_local, halo)
        ok = ok and np.array_equal(plan1.local, exp.local) and plan1.recv_from_periodic == exp.recv_from_periodic \
            and np.array_equal(plan1.send_shift, exp.send_shift)
        moved = int(len(np.setdiff1d(home, plan0.home)))
        # a second repartitioning without motion changes nothing
        h2, x2, s2, l2 = migrate_atoms(t, s.box, world, rank, RLIST, home, xh)
        ok = ok and np.array_equal(h2, home) and np.array_equal(x2, xh) and np.array_equal(s2, send_local) and np.array_equal(l2, halo)
        tot = t.allreduce_sum(torch.tensor([float(len(home))]))
        q.put((rank, bool(ok), moved, int(tot.item()) == s.n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_repartition_migrates_atoms_gloo(world):
    """DD repartitioning (dd_partition_system): after the atoms moved, exchanging leavers with the two neighbours and the
    new halo lists gives every rank exactly the plan it would compute from the global coordinates."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_migrate_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, moved, ok_n in res:
        assert ok and ok_n and moved > 0, (rank, ok, moved, ok_n)


def test_migrate_rejects_long_jumps():
    """an atom that crosses more than one slab between two repartitioning steps is an error, as in the reference"""
    from gmxapi_b200.domdec import LoopbackTransport
    from gmxapi_b200.nblib import InputException
    s = S.water_box(16, 6, 6, seed=5)  # 4.97 nm: 4 slabs of 1.24 nm
    plan = DomainPlan(s.x, s.box, 4, 0, RLIST)
    x = s.x[plan.home].copy()
    x[0, 0] = 3.0  # two slabs away
    with pytest.raises(InputException):
        migrate_atoms(LoopbackTransport(4).endpoint(0), s.box, 4, 0, RLIST, plan.home, x)


# ---- 1-D / 2-D / 3-D decomposition with the half-shell rule (gmxapi_b200/domdec_nd.py) -------------------------------------

def test_half_shell_offsets():
    assert half_shell_offsets((1, 1, 1)) == []
    assert half_shell_offsets((3, 1, 1)) == [(1, 0, 0)]
    assert len(half_shell_offsets((2, 2, 1))) == 4 and len(half_shell_offsets((2, 2, 2))) == 13
    for grid in ((2, 2, 1), (2, 2, 2), (1, 3, 2)):
        offs = half_shell_offsets(grid)
        assert all(tuple(-c for c in o) not in offs for o in offs)  # of o and -o exactly one is in the half shell
        allo = [o for o in np.ndindex(3, 3, 3)]
        assert len(offs) * 2 + 1 == len([o for o in allo if all(grid[d] > 1 or o[d] == 1 for d in range(3))])


@pytest.mark.parametrize("grid", [(2, 2, 1), (2, 2, 2), (3, 2, 1), (1, 2, 3), (4, 1, 1)])
def test_plan_nd_partitions_atoms_and_pairs(grid):
    """Half-shell rule: every atom has one owner, no atom is imported twice, and the union over ranks of home x home and
    home x halo pairs within rlist is every pair of the fully periodic system exactly once."""
    r = 0.45  # small list radius so that small boxes satisfy width >= 2 r with two ranks
    s = S.water_box(8, 6, 6, seed=9)  # 2.49 x 1.86 x 1.86 nm
    x, box = s.x, s.box
    nr = int(np.prod(grid))
    plans = [DomainPlanND(x, box, grid, k, r) for k in range(nr)]
    assert np.array_equal(np.sort(np.concatenate([p.home for p in plans])), np.arange(s.n))
    # reference: all pairs within r, minimum image in every dimension
    d = x[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)
    d -= np.round(d / box) * box
    ref = np.argwhere(np.triu((d ** 2).sum(-1) < r * r, 1))
    ref_keys = np.sort(ref[:, 0].astype(np.int64) * s.n + ref[:, 1])
    got = []
    for p in plans:
        assert len(p.send) == len(p.recv) == len(half_shell_offsets(grid))
        xh = x[p.home].astype(np.float64)
        # along dimensions that are NOT decomposed the kernel applies the periodic images itself: minimum image there
        per = np.array([g == 1 for g in grid])

        def pairs(xa, xb, same):
            dd = xa[:, None, :] - xb[None, :, :]
            dd -= np.where(per, np.round(dd / box) * box, 0.0)
            m = (dd ** 2).sum(-1) < r * r
            if same:
                m = np.triu(m, 1)
            return np.argwhere(m)
        hh = pairs(xh, xh, True)
        got.append(np.stack([p.home[hh[:, 0]], p.home[hh[:, 1]]], 1))
        if p.nhalo:
            hj = pairs(xh, p.halo_x(x).astype(np.float64), False)
            got.append(np.stack([p.home[hj[:, 0]], p.halo[hj[:, 1]]], 1))
    got = np.concatenate(got)
    lo, hi = np.minimum(got[:, 0], got[:, 1]), np.maximum(got[:, 0], got[:, 1])
    keys = np.sort(lo.astype(np.int64) * s.n + hi)
    assert len(keys) == len(np.unique(keys)), "a pair is computed twice"
    assert np.array_equal(keys, ref_keys)
    # what a rank sends is what its peer expects, message by message, in the same order
    for p in plans:
        for k, sd in enumerate(p.send):
            q = plans[sd["rank"]]
            assert q.recv[k]["rank"] == p.rank and np.array_equal(q.recv[k]["ids"], p.home[sd["local"]])
            assert np.array_equal(q.recv[k]["shift"], sd["shift"])


def test_plan_nd_rejects_thin_domains():
    from gmxapi_b200.nblib import InputException
    s = S.water_box(8, 6, 6, seed=9)
    with pytest.raises(InputException):
        DomainPlanND(s.x, s.box, (2, 2, 1), 0, 0.9)  # 0.93 nm wide in y with two ranks: < 2 x 0.9


def _nd_worker(rank, world, port, q, grid):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = S.water_box(8, 6, 6, seed=9)
        r = 0.45
        plan = DomainPlanND(s.x, s.box, grid, rank, r)
        t = TorchDistTransport()
        # coordinates out: pack (restated with torch on the CPU for this host-logic test), exchange, compare with the plan
        x = torch.from_numpy(s.x[plan.local].copy())
        x[plan.nhome:] = float("nan")
        sends = [(x[:plan.nhome][torch.from_numpy(sd["local"]).long()] + torch.from_numpy(sd["shift"].astype(np.float32) * s.box), sd["rank"])
                 for sd in plan.send]
        recvs, off = [], plan.nhome
        for rv in plan.recv:
            recvs.append((x[off:off + len(rv["ids"])], rv["rank"]))
            off += len(rv["ids"])
        t.exchange([(a.contiguous(), d) for a, d in sends], recvs)
        ok_x = bool(np.array_equal(x[plan.nhome:].numpy(), plan.halo_x(s.x)))
        # forces back: every halo atom carries its global index; the owner must get exactly one contribution per sent atom
        f = torch.zeros((plan.nhome + plan.nhalo, 3))
        f[plan.nhome:] = torch.from_numpy(plan.halo.astype(np.float32))[:, None]
        back = [torch.zeros((len(sd["local"]), 3)) for sd in plan.send]
        fs, off = [], plan.nhome
        for rv in plan.recv:
            fs.append((f[off:off + len(rv["ids"])].contiguous(), rv["rank"]))
            off += len(rv["ids"])
        t.exchange(fs, [(b, sd["rank"]) for b, sd in zip(back, plan.send)])
        ok_f = all(np.array_equal(b[:, 0].numpy(), plan.home[sd["local"]].astype(np.float32)) for b, sd in zip(back, plan.send))
        q.put((rank, ok_x, bool(ok_f), plan.nhalo))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [(2, 2, 1), (2, 2, 2)])
def test_halo_exchange_nd_gloo(grid):
    """2 x 2 and 2 x 2 x 2 ranks over gloo: with two ranks along a dimension a neighbour is reached through both faces, so
    several messages travel between the same two ranks and must match in order."""
    world = int(np.prod(grid))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nd_worker, args=(r, world, port, q, grid)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_x, ok_f, nhalo in res:
        assert ok_x and ok_f and nhalo > 0, (rank, ok_x, ok_f, nhalo)


def _migrate_nd_worker(rank, world, port, q, grid):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = S.water_box(8, 6, 6, seed=9)
        r = 0.45
        t = TorchDistTransport()
        plan0 = DomainPlanND(s.x, s.box, grid, rank, r)
        x1 = _moved(s, seed=13, amp=0.3)  # the same displaced coordinates on every rank; atoms cross faces, edges and the box
        home, xh, send_locals, recv_ids = migrate_atoms_nd(t, s.box, grid, rank, r, plan0.home, x1[plan0.home])
        x1w = wrap_into_box(x1, s.box)
        exp = DomainPlanND(x1w, s.box, grid, rank, r)  # what a plan made from the global coordinates says
        ok = np.array_equal(home, exp.home) and np.array_equal(xh, x1w[exp.home])
        ok = ok and all(np.array_equal(a, b["local"]) for a, b in zip(send_locals, exp.send))
        ok = ok and all(np.array_equal(a, b["ids"]) for a, b in zip(recv_ids, exp.recv))
        plan1 = DomainPlanND.from_parts(s.box, grid, rank, r, home, send_locals, recv_ids)
        ok = ok and np.array_equal(plan1.local, exp.local) and np.array_equal(plan1.halo_x(x1w), exp.halo_x(x1w))
        ok = ok and all(a["rank"] == b["rank"] and np.array_equal(a["shift"], b["shift"]) for a, b in zip(plan1.send + plan1.recv, exp.send + exp.recv))
        moved = int(len(np.setdiff1d(home, plan0.home)))
        h2, x2, s2, r2 = migrate_atoms_nd(t, s.box, grid, rank, r, home, xh)  # no motion: nothing changes
        ok = ok and np.array_equal(h2, home) and np.array_equal(x2, xh) and all(np.array_equal(a, b) for a, b in zip(s2 + r2, send_locals + recv_ids))
        tot = t.allreduce_sum(torch.tensor([float(len(home))]))
        q.put((rank, bool(ok), moved, int(tot.item()) == s.n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [(2, 2, 1), (2, 2, 2)])
def test_repartition_nd_migrates_atoms_gloo(grid):
    """repartitioning of the N-D decomposition: leavers go to up to 26 neighbour domains, the half-shell halo lists are
    rebuilt by exchange -- every rank ends with exactly the plan the global coordinates give"""
    world = int(np.prod(grid))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_migrate_nd_worker, args=(r, world, port, q, grid)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, moved, ok_n in res:
        assert ok and ok_n and moved > 0, (rank, ok, moved, ok_n)
