"""Spatial domain decomposition of the nonbonded path over the GPUs of one box: one rank per GPU, atoms
partitioned into slabs along x, halo coordinates sent before and halo forces returned after the non-local
kernel, every step; at pair-search steps atoms that left their slab migrate to the neighbour (repartitioning).

Reference behaviour mirrored (paths relative to /root/reference/src/gromacs):
  zones / eighth shell      domdec/domdec.cpp:133-146: with one decomposed dimension there are two zones, home (0)
                            and the +x neighbour's atoms within the list radius of my upper face (1); i-zone 0
                            interacts with j-zones 0 (half list) and 1 (full list): pairlist.cpp:3876-3916
  dd_move_x                 domdec/domdec.cpp:260-356: send home atoms within rlist of my LOWER face to the -x
                            neighbour, shifted by +box on the periodic edge (:300-318); the receiver appends
                            them after its home atoms
  dd_move_f                 domdec/domdec.cpp:358-460: the forces on those atoms travel back and are ADDED to
                            the sender's home forces; on the periodic edge their sum also goes into the shift
                            force of the +x shift (:426-458)
  GPU halo                  domdec/gpuhaloexchange_impl.cu:77-131 pack / unpack kernels (ours: b200nb_halo_pack_x,
                            b200nb_halo_unpack_f), :403-444 the transfer (ours: stores into the neighbour's peer-memory
                            window from inside the step's kernels, b200nb_dd_step; transport send/recv at search steps)
  repartitioning            domdec/partition.cpp dd_partition_system, domdec/redistribute.cpp dd_redistribute_cg
                            (ours: migrate_atoms + DomainRank.repartition)
  non-local gridding        nbnxm.cpp:77-95 nbnxn_put_on_grid_nonlocal (ours: grid 1 of the C ABI)

The decomposition PLAN (who owns which atom, which atoms are sent) is a pure function of the coordinates at
the pair-search step and is computed by every rank for itself and its two neighbours with the same numpy
code; the per-step coordinate / force halo traffic is real communication.  The transport is pluggable:
torch.distributed (NCCL on GPUs; gloo for the CPU tests of the plan and the exchange pattern) or an
in-process loopback used by the single-GPU parity tests.
"""
import numpy as np

from . import lib as _lib
from .nblib import InputException, configure_interactions

SHIFT_PLUS_X = 5 * (3 * 1 + 1) + 1 + 2  # XYZ2IS(+1, 0, 0), pbcutil/ishift.h:50


# ---------------------------------------------------------------------------------------------------------
# the plan: pure host logic, no device, no communication
# ---------------------------------------------------------------------------------------------------------
class DomainPlan:
    """Slab `rank` of `nranks` along x.

    home        global indices of the atoms this rank owns, ascending
    send_local  positions in `home` of the atoms sent to the -x neighbour (x - lo < rlist)
    send_shift  vector added to them (+box_x on rank 0: they appear beyond the last rank's upper face)
    halo        global indices of the atoms received from the +x neighbour, in arrival order
    """

    def __init__(self, x, box, nranks, rank, rlist):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3)
        box = np.asarray(box, dtype=np.float32).reshape(3)
        if nranks < 1 or not (0 <= rank < nranks):
            raise InputException("bad rank / nranks")
        width = float(box[0]) / nranks
        if nranks > 1 and width < rlist:
            # one pulse only: the halo must come from the nearest neighbour alone (the reference would add
            # pulses, domdec/domdec.cpp:274; outside this implementation's scope)
            raise InputException("domain width %.3f < list radius %.3f: more than one halo pulse needed" % (width, rlist))
        self.nranks, self.rank, self.rlist, self.box = nranks, rank, float(rlist), box
        self.bounds = self.boundaries(box, nranks)
        owner = self.owner_of(x, box, nranks)
        self.lo, self.hi = float(self.bounds[rank]), float(self.bounds[rank + 1])
        self.home = np.nonzero(owner == rank)[0].astype(np.int32)
        self.left, self.right = (rank - 1) % nranks, (rank + 1) % nranks
        if nranks == 1:
            self.send_local = np.zeros(0, np.int32)
            self.halo = np.zeros(0, np.int32)
            self.send_shift = np.zeros(3, np.float32)
            self.recv_from_periodic = False
        else:
            self.send_local = self._send_list(x, self.home, self.bounds[rank], rlist)
            rhome = np.nonzero(owner == self.right)[0].astype(np.int32)
            self.halo = rhome[self._send_list(x, rhome, self.bounds[self.right], rlist)]
            self.send_shift = np.array([box[0] if rank == 0 else 0.0, 0.0, 0.0], np.float32)
            self.recv_from_periodic = (self.right == 0)
        self.nhome, self.nhalo = len(self.home), len(self.halo)
        self.local = np.concatenate([self.home, self.halo]).astype(np.int32)

    @classmethod
    def from_parts(cls, box, nranks, rank, rlist, home, send_local, halo):
        """A plan assembled from what the ranks exchanged at a repartitioning step (migrate_atoms) instead of from the
        global coordinates."""
        p = cls.__new__(cls)
        p.nranks, p.rank, p.rlist = nranks, rank, float(rlist)
        p.box = np.asarray(box, dtype=np.float32).reshape(3)
        p.bounds = cls.boundaries(p.box, nranks)
        p.lo, p.hi = float(p.bounds[rank]), float(p.bounds[rank + 1])
        p.left, p.right = (rank - 1) % nranks, (rank + 1) % nranks
        p.home = np.ascontiguousarray(home, dtype=np.int32)
        p.send_local = np.ascontiguousarray(send_local, dtype=np.int32)
        p.halo = np.ascontiguousarray(halo, dtype=np.int32)
        p.send_shift = np.array([p.box[0] if (rank == 0 and nranks > 1) else 0.0, 0.0, 0.0], np.float32)
        p.recv_from_periodic = nranks > 1 and p.right == 0
        p.nhome, p.nhalo = len(p.home), len(p.halo)
        p.local = np.concatenate([p.home, p.halo]).astype(np.int32)
        return p

    @staticmethod
    def boundaries(box, nranks):
        return (np.arange(nranks + 1, dtype=np.float64) * (float(box[0]) / nranks)).astype(np.float32)

    @staticmethod
    def owner_of(x, box, nranks):
        b = DomainPlan.boundaries(box, nranks)
        o = np.searchsorted(b, x[:, 0], side="right") - 1
        return np.clip(o, 0, nranks - 1).astype(np.int32)

    @staticmethod
    def _send_list(x, home, lo, rlist):
        return np.nonzero(x[home, 0] - np.float32(lo) < np.float32(rlist))[0].astype(np.int32)

    def local_topology(self, types, q, excl_off, excl_idx):
        """types / charges / exclusions of home + halo atoms with exclusions renumbered to LOCAL indices
        (partners that are neither home nor halo on this rank are dropped: they cannot be in range here)."""
        types = np.asarray(types)
        q = np.asarray(q)
        n = len(types)
        g2l = np.full(n, -1, np.int64)
        g2l[self.local] = np.arange(len(self.local))
        cnt = (np.asarray(excl_off[1:]) - np.asarray(excl_off[:-1]))[self.local]
        starts = np.asarray(excl_off)[self.local]
        tot = int(cnt.sum())
        # gather the CSR rows of the local atoms
        row = np.repeat(np.arange(len(self.local)), cnt)
        pos = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(starts, cnt)
        partner = g2l[np.asarray(excl_idx)[pos]]
        keep = partner >= 0
        row, partner = row[keep], partner[keep]
        off = np.zeros(len(self.local) + 1, np.int64)
        np.add.at(off, row + 1, 1)
        off = np.cumsum(off)
        return (types[self.local].astype(np.int32), q[self.local].astype(np.float32), off.astype(np.int32),
                partner.astype(np.int32))

    def halo_x(self, x):
        """What the halo coordinates must be after dd_move_x (used by the tests as the expected value)."""
        h = np.asarray(x, np.float32)[self.halo].copy()
        if self.recv_from_periodic:
            h[:, 0] += self.box[0]
        return h


class DevicePlan(DomainPlan):
    """The plan of a repartitioning step that ran on the GPU: the global indices of the home and halo atoms live on the device
    (home_dev, halo_dev); home / halo / local are downloaded when somebody asks (a caller that keeps per-atom state of its own
    permutes it with plan.home; the per-step path never needs them)."""

    def __init__(self, box, nranks, rank, rlist, home_dev, send_local, halo_dev):
        self.nranks, self.rank, self.rlist = nranks, rank, float(rlist)
        self.box = np.asarray(box, dtype=np.float32).reshape(3)
        self.bounds = self.boundaries(self.box, nranks)
        self.lo, self.hi = float(self.bounds[rank]), float(self.bounds[rank + 1])
        self.left, self.right = (rank - 1) % nranks, (rank + 1) % nranks
        self.home_dev, self.halo_dev = home_dev, halo_dev
        self.send_local = np.ascontiguousarray(send_local, dtype=np.int32)
        self.send_shift = np.array([self.box[0] if (rank == 0 and nranks > 1) else 0.0, 0.0, 0.0], np.float32)
        self.recv_from_periodic = nranks > 1 and self.right == 0
        self.nhome, self.nhalo = int(home_dev.shape[0]), int(halo_dev.shape[0])
        self._home = self._halo = None

    @property
    def home(self):
        if self._home is None:
            self._home = self.home_dev.cpu().numpy().astype(np.int32)
        return self._home

    @property
    def halo(self):
        if self._halo is None:
            self._halo = self.halo_dev.cpu().numpy().astype(np.int32)
        return self._halo

    @property
    def local(self):
        return np.concatenate([self.home, self.halo]).astype(np.int32)


# ---------------------------------------------------------------------------------------------------------
# repartitioning: atoms that left their slab change owner (dd_partition_system, domdec/partition.cpp: dd_redistribute_cg
# domdec/redistribute.cpp moves them to the neighbour cell, setup_dd_communication rebuilds the halo send lists, and
# GpuHaloExchange::reinitHalo (domdec/gpuhaloexchange_impl.cu:133-213) takes the new index maps)
# ---------------------------------------------------------------------------------------------------------
def wrap_into_box(x, box):
    """put_atoms_in_box for a rectangular box (pbcutil/pbc.cpp): x - floor(x / box) * box, in float32 like the reference."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3).copy()
    box = np.asarray(box, dtype=np.float32).reshape(3)
    for d in range(3):
        while True:
            lo = x[:, d] < 0
            hi = x[:, d] >= box[d]
            if not (lo.any() or hi.any()):
                break
            x[lo, d] += box[d]
            x[hi, d] -= box[d]
    return x


def migrate_atoms(t, box, nranks, rank, rlist, home, x_home, to_tensor=None):
    """One repartitioning step of the x-slab decomposition, collective over the ranks of transport `t`.

    home / x_home: the global indices and CURRENT coordinates of the atoms this rank owned so far.  Atoms whose wrapped
    x coordinate now lies in a neighbour's slab are handed to that neighbour (one slab at most: the reference makes the
    same assumption per repartitioning, redistribute.cpp "moved more than one cell"); then every rank tells its -x
    neighbour which of its atoms lie within rlist of its lower face (the new halo).  Returns
    (home_new ascending, x_home_new, send_local, halo) -- exactly what DomainPlan computes from global coordinates.
    `to_tensor`: numpy int32 array -> tensor the transport can move (CUDA for NCCL, CPU for gloo / loopback)."""
    import torch
    if to_tensor is None:
        to_tensor = torch.from_numpy
    box = np.asarray(box, dtype=np.float32).reshape(3)
    home = np.ascontiguousarray(home, dtype=np.int32)
    x = wrap_into_box(x_home, box)
    left, right = (rank - 1) % nranks, (rank + 1) % nranks
    if nranks == 1:
        order = np.argsort(home, kind="stable")
        return home[order], x[order], np.zeros(0, np.int32), np.zeros(0, np.int32)
    owner = DomainPlan.owner_of(x, box, nranks)
    stay = owner == rank
    go_l = (owner == left) & ~stay
    go_r = (owner == right) & ~stay & ~go_l if left != right else np.zeros_like(stay)
    if left == right:
        # two ranks: both faces lead to the same neighbour; everything that leaves goes "left" (the receiver does not care)
        go_l = ~stay
    if not np.all(stay | go_l | go_r):
        raise InputException("an atom moved more than one domain between two repartitioning steps")

    def pack(mask):
        m = np.nonzero(mask)[0]
        buf = np.empty((len(m), 4), np.int32)
        buf[:, 0] = home[m]
        buf[:, 1:] = x[m].view(np.int32)
        return buf

    out_l, out_r = pack(go_l), pack(go_r)
    counts = t.allgather_object(dict(to_left=len(out_l), to_right=len(out_r)))
    # what arrives from the right neighbour is what it sends to ITS left, and vice versa
    in_r = np.empty((counts[right]["to_left"], 4), np.int32)
    in_l = np.empty((counts[left]["to_right"], 4), np.int32)

    def exchange(send_np, dst, recv_np, src):
        send = to_tensor(np.ascontiguousarray(send_np)) if len(send_np) else None
        recv = to_tensor(recv_np) if len(recv_np) else None
        t.sendrecv(send, dst, recv, src)
        if recv is not None:
            recv_np[...] = recv.cpu().numpy()

    exchange(out_l, left, in_r, right)
    exchange(out_r, right, in_l, left)
    arrived = np.concatenate([in_r, in_l]) if (len(in_r) or len(in_l)) else np.zeros((0, 4), np.int32)
    ids = np.concatenate([home[stay], arrived[:, 0]]).astype(np.int32)
    xs = np.concatenate([x[stay], arrived[:, 1:].copy().view(np.float32)]).astype(np.float32)
    order = np.argsort(ids, kind="stable")
    home_new, x_new = ids[order], np.ascontiguousarray(xs[order])
    bounds = DomainPlan.boundaries(box, nranks)
    if len(home_new) and not np.array_equal(DomainPlan.owner_of(x_new, box, nranks), np.full(len(home_new), rank, np.int32)):
        raise InputException("repartitioning left an atom outside its new owner's slab")
    send_local = DomainPlan._send_list(x_new, np.arange(len(home_new)), bounds[rank], rlist)
    nsend = t.allgather_object(int(len(send_local)))
    halo = np.empty(nsend[right], np.int32)
    exchange(home_new[send_local], left, halo, right)
    return home_new, x_new, send_local, halo


# ---------------------------------------------------------------------------------------------------------
# transports
# ---------------------------------------------------------------------------------------------------------
class TorchDistTransport:
    """Halo transfers over torch.distributed: NCCL send/recv on NVLink for CUDA tensors (one group per exchange,
    issued on the context's stream), gloo for the CPU tests."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.nranks = dist.get_rank(group), dist.get_world_size(group)

    def sendrecv(self, send, dst, recv, src):
        dist = self.dist
        ops = []
        if send is not None and send.numel():
            ops.append(dist.P2POp(dist.isend, send, dst, self.group))
        if recv is not None and recv.numel():
            ops.append(dist.P2POp(dist.irecv, recv, src, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()  # CUDA: orders the current stream after the transfer, does not block the host

    def exchange(self, sends, recvs):
        """Several messages at once (domdec_nd: one per half-shell neighbour): sends = [(tensor, dst)], recvs = [(tensor, src)];
        messages between the same two ranks are matched in list order."""
        dist = self.dist
        ops = [dist.P2POp(dist.isend, t, dst, self.group) for t, dst in sends if t is not None and t.numel()]
        ops += [dist.P2POp(dist.irecv, t, src, self.group) for t, src in recvs if t is not None and t.numel()]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, group=self.group)
        return t

    def barrier(self):
        self.dist.barrier(self.group)

    def allgather_object(self, obj):
        out = [None] * self.nranks
        self.dist.all_gather_object(out, obj, group=self.group)
        return out


class LoopbackTransport:
    """All ranks live in one process as threads (single-GPU parity tests): a queue per (src, dst) pair."""

    def __init__(self, nranks):
        import queue
        self.nranks = nranks
        self.q = {(s, d): queue.Queue() for s in range(nranks) for d in range(nranks)}
        import threading
        self._bar = threading.Barrier(nranks)
        self._objs = [None] * nranks

    def endpoint(self, rank):
        return _LoopbackEndpoint(self, rank)


class _LoopbackEndpoint:
    def __init__(self, hub, rank):
        self.hub, self.rank, self.nranks = hub, rank, hub.nranks
        self.sync = None  # set by DomainRank: synchronises the sender's stream before handing a buffer over

    def sendrecv(self, send, dst, recv, src):
        if send is not None and send.numel():
            if self.sync:
                self.sync()
            self.hub.q[(self.rank, dst)].put(send.clone())
            if self.sync:
                self.sync()
        if recv is not None and recv.numel():
            buf = self.hub.q[(src, self.rank)].get(timeout=120)
            recv.copy_(buf)
            if self.sync:
                self.sync()  # the copy must have run before `buf` returns to the sender's stream pool and is rewritten
            del buf

    def exchange(self, sends, recvs):
        """all sends are queued first (the queues are unbounded), then the receives are taken in list order"""
        live = [(t, dst) for t, dst in sends if t is not None and t.numel()]
        if live and self.sync:
            self.sync()
        for t, dst in live:
            self.hub.q[(self.rank, dst)].put(t.clone())
        if live and self.sync:
            self.sync()
        got = False
        for t, src in recvs:
            if t is not None and t.numel():
                buf = self.hub.q[(src, self.rank)].get(timeout=120)
                t.copy_(buf)
                got = True
                del buf
        if got and self.sync:
            self.sync()

    def allreduce_sum(self, t):
        raise NotImplementedError

    def barrier(self):
        self.hub._bar.wait()

    def allgather_object(self, obj):
        self.hub._objs[self.rank] = obj
        self.hub._bar.wait()
        out = list(self.hub._objs)
        self.hub._bar.wait()
        return out


# ---------------------------------------------------------------------------------------------------------
# one rank of the decomposed calculation
# ---------------------------------------------------------------------------------------------------------
class DomainRank:
    """The nonbonded calculation of one rank: its NbnxmGpu context with a home grid (0) and a halo grid (1), the
    device buffers of the halo exchange, and the per-step schedule of do_force() (mdlib/sim_util.cpp:1388-1902
    restricted to the nonbonded path):

        x home -> grid order + clear | push halo x | local kernel | wait halo x -> grid order | non-local kernel |
        push halo f | wait halo f, add (+ shift force on the edge), forces -> atom order

    Per step everything is ONE library call (b200nb_dd_step): the halos move through peer-memory windows written by
    the neighbours' kernels over NVLink (CUDA IPC), flags instead of host synchronisation, no collective.  The
    transport (NCCL send/recv, or the in-process loopback of the tests) is used at set-up (window handles) and on
    pair-search steps, where the halo composition changes.  `use_windows=False` keeps the transport on the per-step
    path as well (the first version of this file; kept for comparison runs).
    """

    def __init__(self, system, options, transport, rank=None, nranks=None, device=0, use_windows=True):
        import torch
        if np.any(np.asarray(getattr(system, "box_offdiag", np.zeros(3))) != 0):
            raise InputException("domain decomposition of a triclinic cell is not supported")
        self.torch = torch
        self.t = transport
        self.rank = transport.rank if rank is None else rank
        self.nranks = transport.nranks if nranks is None else nranks
        self.options = options
        rc = float(options.pairlistCutoff)
        self.rlist = float(options.rlistOuter or rc)
        self.plan = DomainPlan(system.x, system.box, self.nranks, self.rank, self.rlist)
        p = self.plan
        self.nb = _lib.NbnxmGpu(device)
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)  # torch owns it: its allocator's per-stream pools outlive nb
        self.nb.set_stream(self.stream.cuda_stream)
        if hasattr(transport, "sync"):
            transport.sync = self.nb.synchronize
        configure_interactions(self.nb, system.nbfp, options, self.rlist)
        # the topology is global and replicated on every rank (as the reference keeps gmx_mtop_t); ownership is what moves
        self.topology = (np.asarray(system.types), np.asarray(system.q), np.asarray(system.excl_off), np.asarray(system.excl_idx))
        self.nb.set_box(system.box, pbc=(0 if self.nranks > 1 else 1, 1, 1))
        self.fshift_halo = np.zeros(3, np.float64)
        self.use_windows = bool(use_windows) and self.nranks > 1
        self.max_halo = self.max_send = 0
        self._setup_local(np.ascontiguousarray(system.x[p.home]), first=True)

    def _setup_local(self, x_home, first=False, atoms_installed=False):
        """(Re)build everything that depends on which atoms this rank owns and receives: local topology, device buffers,
        grids, pair list, halo plan.  atoms_installed: the local topology is already in the context (built on the device by
        b200nb_dd_set_local_atoms) and x_home is a device tensor."""
        torch = self.torch
        p = self.plan
        if not atoms_installed:
            types, q, eo, ei = p.local_topology(*self.topology)
            self.nb.set_atoms(types, q, eo, ei)
        self.nlocal = p.nhome + p.nhalo
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.x = torch.zeros((self.nlocal, 3), dtype=torch.float32, device=self.device)
            self.f = torch.zeros((self.nlocal, 3), dtype=torch.float32, device=self.device)
            self.send_idx = torch.from_numpy(p.send_local).to(self.device)
            self.send_buf = torch.zeros((len(p.send_local), 3), dtype=torch.float32, device=self.device)
            self.recv_f = torch.zeros((len(p.send_local), 3), dtype=torch.float32, device=self.device)
            if isinstance(x_home, torch.Tensor):
                self.x[:p.nhome].copy_(x_home)
            else:
                self.x[:p.nhome].copy_(torch.from_numpy(np.ascontiguousarray(x_home, dtype=np.float32)))
        self.nb.synchronize()
        if self.use_windows:
            if first:
                self._open_windows()
            elif p.nhalo > self.max_halo or len(p.send_local) > self.max_send:
                raise InputException("repartition: halo of %d / %d atoms exceeds the window capacity %d / %d"
                                     % (p.nhalo, len(p.send_local), self.max_halo, self.max_send))
        self.search(first=first)

    def repartition(self, x_home=None, on_device=True):
        """Pair-search step WITH atom migration (dd_partition_system, domdec/partition.cpp): atoms whose current
        coordinates (self.x[:nhome], on the device) left this rank's slab change owner, coordinates are wrapped into the
        box, the halo send lists are rebuilt from the new ownership, then grids and pair list are rebuilt.  Collective:
        every rank calls it at the same step.  Returns the new plan (plan.home = global indices now owned, ascending;
        a caller integrating the equations of motion moves its velocities etc. with the same map)."""
        torch = self.torch
        p = self.plan
        self.nb.synchronize()
        if on_device:
            return self._repartition_device(x_home)
        # ---- host variant (numpy; what the gloo tests of the plan logic run, and the cross-check of the device variant) ----
        # the current coordinates of the home atoms: the caller's (x_home, e.g. after an integration step), else those of the
        # last step -- the caller's pinned buffer when the step ran on it in place, self.x otherwise
        xc = getattr(self, "_x_host_current", None)
        if x_home is not None:
            x_home = np.ascontiguousarray(x_home.numpy() if hasattr(x_home, "numpy") else x_home, dtype=np.float32).reshape(-1, 3)
            if len(x_home) != p.nhome:
                raise InputException("repartition: x_home must hold the %d home atoms" % p.nhome)
        elif xc is not None and len(xc) == p.nhome:
            x_home = xc.numpy().copy()
        else:
            x_home = self.x[:p.nhome].cpu().numpy()
        self._x_host_current = None
        to_dev = lambda a: torch.from_numpy(a).to(self.device)
        home, x_new, send_local, halo = migrate_atoms(self.t, p.box, self.nranks, self.rank, self.rlist, p.home, x_home,
                                                      to_tensor=to_dev)
        self.plan = DomainPlan.from_parts(p.box, self.nranks, self.rank, self.rlist, home, send_local, halo)
        self._setup_local(x_new)
        return self.plan

    def _repartition_device(self, x_home):
        """repartition() with the coordinate-dependent work on the GPU (csrc/dd_partition.cu): wrap + ownership, stable compaction
        of stayers / leavers, message packing, merge of stayers and arrivals into the new home set, halo selection, local
        topology with exclusions renumbered through a device-resident global -> local look-up.  Coordinates stay on the device;
        the host sees the counts (a dozen ints) and the send list that b200nb_dd_set_plan takes.  Messages between the ranks
        travel as device tensors through the transport (NCCL send/recv; the loopback of the single-GPU tests)."""
        torch = self.torch
        p, nb, t = self.plan, self.nb, self.t
        dev = self.device
        i32 = dict(dtype=torch.int32, device=dev)
        with torch.cuda.device(dev), torch.cuda.stream(self.stream):
            if not getattr(self, "_global_topology_on_device", False):
                nb.dd_set_global_topology(*self.topology)
                self._global_topology_on_device = True
            xc = getattr(self, "_x_host_current", None)
            if x_home is not None:
                xt = x_home if isinstance(x_home, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_home, dtype=np.float32))
                if tuple(xt.shape) != (p.nhome, 3):
                    raise InputException("repartition: x_home must hold the %d home atoms" % p.nhome)
                x = xt.to(dev, dtype=torch.float32, copy=True).contiguous()
            elif xc is not None and len(xc) == p.nhome:
                x = xc.to(dev, copy=True)
            else:
                x = self.x[:p.nhome].clone()
            self._x_host_current = None
            gid = getattr(p, "home_dev", None)
            if gid is None:
                gid = torch.from_numpy(np.ascontiguousarray(p.home, dtype=np.int32)).to(dev)
            n = p.nhome
            bounds = DomainPlan.boundaries(p.box, self.nranks)
            code = torch.empty(max(n, 1), **i32)
            idx = torch.empty(max(n, 1), **i32)
            nb.dd_wrap_classify(x.data_ptr(), n, p.box, bounds, self.nranks, self.rank, code.data_ptr())
            nstay, nl, nr, nlost = nb.dd_partition_indices(code.data_ptr(), n, 4, idx.data_ptr())
            if self.nranks == 1:
                nstay, nl, nr, nlost = n, 0, 0, 0
            if nlost:
                raise InputException("an atom moved more than one domain between two repartitioning steps")
            out_l = torch.empty((nl, 4), **i32)
            out_r = torch.empty((nr, 4), **i32)
            nb.dd_pack_atoms(idx.data_ptr() + 4 * nstay, nl, gid.data_ptr(), x.data_ptr(), out_l.data_ptr())
            nb.dd_pack_atoms(idx.data_ptr() + 4 * (nstay + nl), nr, gid.data_ptr(), x.data_ptr(), out_r.data_ptr())
            if self.nranks > 1:
                counts = t.allgather_object(dict(to_left=nl, to_right=nr))
                # what arrives from the right neighbour is what it sends to ITS left, and vice versa
                in_r = torch.empty((counts[p.right]["to_left"], 4), **i32)
                in_l = torch.empty((counts[p.left]["to_right"], 4), **i32)
                t.sendrecv(out_l, p.left, in_r, p.right)
                t.sendrecv(out_r, p.right, in_l, p.left)
                arrived = torch.cat([in_r, in_l]).contiguous()
            else:
                arrived = torch.empty((0, 4), **i32)
            m = int(arrived.shape[0])
            n_new = nstay + m
            gid_new = torch.empty(max(n_new, 1), **i32)[:n_new]
            x_new = torch.empty((max(n_new, 1), 3), dtype=torch.float32, device=dev)[:n_new]
            nb.dd_merge_home(idx.data_ptr(), nstay, gid.data_ptr(), x.data_ptr(), arrived.data_ptr() if m else 0, m,
                             gid_new.data_ptr(), x_new.data_ptr())
            # every atom of the new home set must lie in this rank's slab (the host variant's owner_of check), and the halo of the
            # -x neighbour = the home atoms within rlist of the lower face: one more classification + partition each
            code2 = torch.empty(max(n_new, 1), **i32)
            idx2 = torch.empty(max(n_new, 1), **i32)
            if self.nranks > 1:
                nb.dd_wrap_classify(x_new.data_ptr(), n_new, p.box, bounds, self.nranks, self.rank, code2.data_ptr())
                if nb.dd_partition_indices(code2.data_ptr(), n_new, 4, idx2.data_ptr())[0] != n_new:
                    raise InputException("repartitioning left an atom outside its new owner's slab")
                nb.dd_select_lower_face(x_new.data_ptr(), n_new, float(bounds[self.rank]), self.rlist, code2.data_ptr())
                nkeep, nsend = nb.dd_partition_indices(code2.data_ptr(), n_new, 2, idx2.data_ptr())
                send_local_dev = idx2[nkeep:nkeep + nsend]
                send_gid = torch.empty(max(nsend, 1), **i32)[:nsend]
                nb.dd_gather_int(idx2.data_ptr() + 4 * nkeep, nsend, gid_new.data_ptr(), send_gid.data_ptr())
                nsends = t.allgather_object(int(nsend))
                halo_gid = torch.empty(max(nsends[p.right], 1), **i32)[:nsends[p.right]]
                t.sendrecv(send_gid, p.left, halo_gid, p.right)
                send_local = send_local_dev.cpu().numpy().astype(np.int32)  # b200nb_dd_set_plan takes it from the host: a few thousand ints
            else:
                halo_gid = torch.empty(0, **i32)
                send_local = np.zeros(0, np.int32)
            local_gid = torch.cat([gid_new, halo_gid]).contiguous()
            nb.dd_set_local_atoms(local_gid.data_ptr(), int(local_gid.shape[0]))
        self.plan = DevicePlan(p.box, self.nranks, self.rank, self.rlist, gid_new, send_local, halo_gid)
        self._setup_local(x_new, atoms_installed=True)
        return self.plan

    # -- peer-memory halo windows: created once, sized with slack for later search steps -----------------------------------
    def _open_windows(self):
        import os
        p = self.plan
        max_halo = int(1.5 * p.nhalo) + 4096
        max_send = int(1.5 * len(p.send_local)) + 4096
        self.max_halo, self.max_send = max_halo, max_send
        handle, ptr = self.nb.dd_create_window(max_halo, max_send)
        info = self.t.allgather_object(dict(pid=os.getpid(), handle=handle, ptr=ptr, max_halo=max_halo))
        for side, peer in ((0, p.left), (1, p.right)):
            pi = info[peer]
            if pi["pid"] == os.getpid():
                self.nb.dd_open_peer(side, window_ptr=pi["ptr"], peer_max_halo=pi["max_halo"])
            else:
                self.nb.dd_open_peer(side, ipc_handle=pi["handle"], peer_max_halo=pi["max_halo"])
        self.t.barrier()

    # -- pair search step: put_on_grid (home, then halo) + constructPairlist: mdlib/sim_util.cpp:1316-1366, 1454 ------
    def search(self, first=False):
        p = self.plan
        box = p.box
        self._halo_x()
        self.nb.synchronize()
        lower = np.array([p.lo if self.nranks > 1 else 0.0, 0, 0], np.float32)
        upper = np.array([p.hi if self.nranks > 1 else box[0], box[1], box[2]], np.float32)
        self.nb.put_on_grid(self.x.data_ptr(), lower, upper, 0, 0, p.nhome, on_device=True)
        if p.nhalo:
            hl = np.array([upper[0], 0, 0], np.float32)
            hu = np.array([upper[0] + self.rlist, box[1], box[2]], np.float32)
            self.nb.put_on_grid(self.x.data_ptr(), hl, hu, 1, p.nhome, self.nlocal, on_device=True)
        self.nb.build_pairlist()
        if self.use_windows:
            # the forces this rank computes on halo atoms that arrived across the periodic edge (the last slab's halo is
            # rank 0's lower boundary, shifted by +box) also enter the +x shift force (domdec/domdec.cpp:426-458)
            edge = SHIFT_PLUS_X if p.recv_from_periodic else -1
            self.nb.dd_set_plan(p.nhome, p.nhalo, p.send_local, p.send_shift, edge)
            self.t.barrier()  # every rank has its plan before anybody steps

    # -- dd_move_x ------------------------------------------------------------------------------------------------------
    def _halo_x(self):
        p = self.plan
        if self.nranks == 1:
            return
        torch = self.torch
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.nb.halo_pack_x(self.x.data_ptr(), self.send_idx.data_ptr(), len(p.send_local), p.send_shift,
                                self.send_buf.data_ptr())
            self.t.sendrecv(self.send_buf, p.left, self.x[p.nhome:], p.right)

    # -- dd_move_f ------------------------------------------------------------------------------------------------------
    def _halo_f(self, virial):
        p = self.plan
        if self.nranks == 1:
            return
        torch = self.torch
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.t.sendrecv(self.f[p.nhome:], p.right, self.recv_f, p.left)
            self.nb.halo_unpack_f(self.f.data_ptr(), self.send_idx.data_ptr(), len(p.send_local), self.recv_f.data_ptr())
            if virial and self.rank == 0:
                # domdec/domdec.cpp:426-458: forces that crossed the periodic edge also enter the shift forces
                self.fshift_halo = self.recv_f.sum(0, dtype=torch.float64).cpu().numpy()

    def step(self, flags=0):
        """One nonbonded step on coordinates already in self.x[:nhome] (device). Leaves forces in self.f[:nhome]."""
        p = self.plan
        if self.use_windows:
            self.nb.dd_step(self.x.data_ptr(), self.f.data_ptr(), flags)
            return
        self.nb.set_x(self.x.data_ptr(), on_device=True, atom_begin=0, atom_end=p.nhome)
        self.nb.clear_outputs()
        self.nb.launch_force(0, flags)
        if p.nhalo or self.nranks > 1:
            self._halo_x()
            if p.nhalo:
                self.nb.set_x(self.x.data_ptr(), on_device=True, atom_begin=p.nhome, atom_end=self.nlocal)
                self.nb.launch_force(1, flags)
        self.nb.get_f(self.f.data_ptr(), on_device=True)
        self._halo_f(bool(flags & _lib.FLAG_VIRIAL))

    def compute(self, x_home_host, flags=0, f_home_host=None):
        """The public per-rank call: host coordinates of the home atoms in, host forces of the home atoms out
        (H2D and D2H inside). Returns (f_home, fshift[45,3], e_lj, e_el) with this rank's share of the sums."""
        torch = self.torch
        p = self.plan
        xh = x_home_host if isinstance(x_home_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_home_host, np.float32))
        if f_home_host is None:
            f_home_host = torch.empty((p.nhome, 3), dtype=torch.float32).pin_memory()
        fh = f_home_host if isinstance(f_home_host, torch.Tensor) else torch.from_numpy(f_home_host)
        if self.use_windows and xh.is_pinned() and fh.is_pinned():
            # the kernels read the pinned coordinates and write the pinned forces in place (no staging copy); repartition()
            # takes the current coordinates from this buffer (self.x is not refreshed on this path)
            self.nb.dd_step(xh.data_ptr(), fh.data_ptr(), flags)
            self._x_host_current = xh
        else:
            self._x_host_current = None
            with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
                self.x[:p.nhome].copy_(xh, non_blocking=True)
                self.step(flags)
                fh.copy_(self.f[:p.nhome], non_blocking=True)
        self.nb.synchronize()
        if self.use_windows:
            self.nb.dd_status()
        fs = np.zeros((_lib.SHIFTS, 3), np.float32)
        elj = eel = 0.0
        if flags:
            fs, elj, eel = self.nb.get_outputs()
            if self.rank == 0 and self.nranks > 1 and not self.use_windows:
                fs[SHIFT_PLUS_X] += self.fshift_halo.astype(np.float32)
        return fh, fs, elj, eel

    def pair_count(self, r):
        return self.nb.pair_count(r)

    def close(self):
        self.nb.synchronize()
        self.nb.close()
        self.x = self.f = self.send_idx = self.send_buf = self.recv_f = None
