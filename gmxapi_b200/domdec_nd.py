"""Domain decomposition of the nonbonded path in one, two or three dimensions (e.g. 2 x 2 x 2 ranks for the 8 GPUs of a box).

The slab decomposition of `domdec.py` (eighth shell with one decomposed dimension, halos through peer-memory windows) is
the fast path; this module is the general one.  It assigns pairs with the HALF-SHELL rule instead of the reference's eighth
shell (domdec/domdec.cpp:133-146): rank A imports the atoms near its boundary from the 13 (3-D), 4 (2-D) or 1 (1-D)
neighbours B = A + o whose offset o is lexicographically positive (first non-zero component +1) and computes
home x home (half list) and home x halo (full list) -- every pair of atoms of two different domains is then computed exactly
once, on the rank that sees the other domain at a positive offset, and there are NO halo x halo pairs, so the two grids
(home, halo) and the two list kinds the library has are all that is needed.  The price is a halo about twice the eighth
shell's, and direct exchanges with every neighbour instead of three forwarding pulses -- which is how an NVSwitch box wants
it anyway (every GPU reaches every other at full bandwidth).

Reference behaviour mirrored (paths relative to /root/reference/src/gromacs):
  dd_move_x / dd_move_f        domdec/domdec.cpp:260-460: coordinates of boundary atoms out (shifted by the box vector when they
                               cross a periodic edge, :300-318), forces on them back and added, their sum into the shift force
                               of that shift for the virial (:426-458)
  non-local gridding / search  nbnxm.cpp:77-95, pairlist.cpp:3876-3916: halo atoms are gridded as the second grid, periodic
                               images are switched off along decomposed dimensions
  pack / unpack                domdec/gpuhaloexchange_impl.cu:77-131 (ours: b200nb_halo_pack_x / b200nb_halo_unpack_f)
The per-step exchange runs through the same peer-memory windows as the slab path (b200nb_dd_set_links: one LINK per half-shell
offset, up to 13; the kernels of the step write the halo coordinates / forces straight into the neighbours' windows over NVLink
and wait on per-link flags); `use_windows=False` keeps the transport (NCCL send/recv, or the in-process loopback) on the
per-step path with separate pack / unpack kernels.  The pair-search steps always use the transport.
"""
import itertools

import numpy as np

from . import lib as _lib
from .nblib import InputException, configure_interactions


def shift_index(k):
    """XYZ2IS, pbcutil/ishift.h:50"""
    return 5 * (3 * (int(k[2]) + 1) + (int(k[1]) + 1)) + int(k[0]) + 2


def half_shell_offsets(grid):
    """Offsets o in {-1,0,1}^3, non-zero only along decomposed dimensions, whose first non-zero component is +1."""
    rng = [(-1, 0, 1) if n > 1 else (0,) for n in grid]
    out = []
    for o in itertools.product(*rng):
        nz = [c for c in o if c != 0]
        if nz and nz[0] > 0:
            out.append(tuple(o))
    return out


class DomainPlanND:
    """Rank `rank` of a grid of nx x ny x nz domains (rank = (ix * ny + iy) * nz + iz), equal widths.

    home    global indices owned, ascending
    recv    list of dict(rank, offset, shift[3], ids): what arrives from neighbour `rank` seen at `offset`; `shift` (in box
            vectors, -1/0/+1 per dimension) is what the SENDER adds so the atoms appear next to this domain
    send    list of dict(rank, offset, shift[3], local): positions in `home` of the atoms this rank sends to `rank`, which sees
            this domain at `offset`
    halo    concatenation of the recv ids, in recv order (a global index can appear once only: domains of width >= 2 rlist
            along a dimension with two ranks, see below)
    """

    def __init__(self, x, box, grid, rank, rlist):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 3)
        self._geometry(box, grid, rank, rlist)
        owner = self.owner_of(x, self.box, self.grid)
        self.home = np.nonzero(owner == self.rank)[0].astype(np.int32)
        self.recv, self.send = [], []
        for o in self.offsets:
            # what I receive from the neighbour at +o: its atoms near the boundary it shares with me
            nb, shift, dst, sshift = self._peers(o)
            nb_home = np.nonzero(owner == nb)[0].astype(np.int32)
            sel = self.boundary_atoms(x[nb_home], self.coords_of(nb, self.grid), o)
            self.recv.append(dict(rank=nb, offset=o, shift=shift, ids=nb_home[sel]))
            # what I send to the rank that sees me at +o
            self.send.append(dict(rank=dst, offset=o, shift=sshift,
                                  local=self.boundary_atoms(x[self.home], self.coords, o).astype(np.int32)))
        self._finish()

    @classmethod
    def from_parts(cls, box, grid, rank, rlist, home, send_locals, recv_ids):
        """A plan assembled from what the ranks exchanged at a repartitioning step (migrate_atoms_nd): the new home set and,
        per half-shell offset, the positions sent and the global indices received."""
        p = cls.__new__(cls)
        p._geometry(box, grid, rank, rlist)
        p.home = np.ascontiguousarray(home, dtype=np.int32)
        p.recv, p.send = [], []
        for k, o in enumerate(p.offsets):
            nb, shift, dst, sshift = p._peers(o)
            p.recv.append(dict(rank=nb, offset=o, shift=shift, ids=np.ascontiguousarray(recv_ids[k], dtype=np.int32)))
            p.send.append(dict(rank=dst, offset=o, shift=sshift, local=np.ascontiguousarray(send_locals[k], dtype=np.int32)))
        p._finish()
        return p

    def _geometry(self, box, grid, rank, rlist):
        self.box = np.asarray(box, dtype=np.float32).reshape(3)
        self.grid = tuple(int(g) for g in grid)
        self.nranks = int(np.prod(self.grid))
        if len(self.grid) != 3 or min(self.grid) < 1 or not (0 <= rank < self.nranks):
            raise InputException("bad grid / rank")
        self.rank, self.rlist = int(rank), float(rlist)
        self.width = self.box.astype(np.float64) / np.array(self.grid)
        for d in range(3):
            if self.grid[d] > 1 and self.width[d] < rlist:
                raise InputException("domain width %.3f < list radius %.3f along dimension %d" % (self.width[d], rlist, d))
            if self.grid[d] == 2 and self.width[d] < 2 * rlist:
                # both faces lead to the same neighbour: an atom within rlist of both would be imported as two images
                raise InputException("two domains along dimension %d need a width of at least 2 x the list radius" % d)
        self.coords = self.coords_of(self.rank, self.grid)
        self.lo = np.array([self.bound(d, self.coords[d]) for d in range(3)], np.float32)
        self.hi = np.array([self.bound(d, self.coords[d] + 1) for d in range(3)], np.float32)
        self.offsets = half_shell_offsets(self.grid)

    def _peers(self, o):
        """(rank I receive from at +o, shift of its atoms; rank that sees me at +o, shift of my atoms there)"""
        nb, shift = self.neighbour(self.coords, o, +1)
        dst, _ = self.neighbour(self.coords, o, -1)
        _, sshift = self.neighbour(self.coords_of(dst, self.grid), o, +1)
        return nb, shift, dst, sshift

    def _finish(self):
        self.halo = (np.concatenate([r["ids"] for r in self.recv]) if self.recv else np.zeros(0, np.int32)).astype(np.int32)
        if len(np.unique(self.halo)) != len(self.halo):
            raise InputException("an atom would be imported twice (domain too thin for its number of ranks)")
        self.nhome, self.nhalo = len(self.home), len(self.halo)
        self.local = np.concatenate([self.home, self.halo]).astype(np.int32)

    # -- geometry ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def coords_of(rank, grid):
        iz = rank % grid[2]
        iy = (rank // grid[2]) % grid[1]
        ix = rank // (grid[1] * grid[2])
        return (ix, iy, iz)

    @staticmethod
    def rank_of(coords, grid):
        return (coords[0] * grid[1] + coords[1]) * grid[2] + coords[2]

    def bound(self, d, i):
        return np.float32(i * (float(self.box[d]) / self.grid[d]))

    @staticmethod
    def owner_of(x, box, grid):
        idx = []
        for d in range(3):
            b = (np.arange(grid[d] + 1, dtype=np.float64) * (float(box[d]) / grid[d])).astype(np.float32)
            idx.append(np.clip(np.searchsorted(b, x[:, d], side="right") - 1, 0, grid[d] - 1))
        return ((idx[0] * grid[1] + idx[1]) * grid[2] + idx[2]).astype(np.int32)

    def neighbour(self, coords, o, sign):
        """Rank at coords + sign*o (periodic) and the shift, in box vectors, under which ITS atoms appear next to `coords`."""
        c, shift = [], []
        for d in range(3):
            v = coords[d] + sign * o[d]
            k = 0
            if v >= self.grid[d]:
                v -= self.grid[d]
                k = +1
            elif v < 0:
                v += self.grid[d]
                k = -1
            c.append(v)
            shift.append(k)
        return self.rank_of(c, self.grid), np.array(shift, np.int32)

    def boundary_atoms(self, xs, coords, o):
        """Positions (in xs) of the atoms of domain `coords` that a domain seeing it at offset +o needs: within rlist of the
        lower face along dimensions with o = +1, of the upper face where o = -1 (slab criteria: a superset of the sphere)."""
        m = np.ones(len(xs), bool)
        r = np.float32(self.rlist)
        for d in range(3):
            if o[d] == +1:
                m &= xs[:, d] - self.bound(d, coords[d]) < r
            elif o[d] == -1:
                m &= self.bound(d, coords[d] + 1) - xs[:, d] <= r
        return np.nonzero(m)[0]

    # -- derived data -------------------------------------------------------------------------------------------------------
    def halo_x(self, x):
        """What the halo coordinates must be after the coordinate exchange (expected values for the tests)."""
        x = np.asarray(x, np.float32)
        parts = [x[r["ids"]] + (r["shift"].astype(np.float32) * self.box) for r in self.recv]
        return np.concatenate(parts).astype(np.float32) if parts else np.zeros((0, 3), np.float32)

    def local_topology(self, types, q, excl_off, excl_idx):
        """types / charges / exclusions of home + halo atoms, exclusions renumbered to local indices (as DomainPlan)."""
        from .domdec import DomainPlan
        return DomainPlan.local_topology(self, types, q, excl_off, excl_idx)

    def halo_bounds(self):
        """Bounding region of the halo grid: the domain grown by rlist on the faces a half-shell neighbour can sit behind."""
        lo, hi = self.lo.copy(), self.hi.copy()
        for d in range(3):
            if any(o[d] == +1 for o in self.offsets):
                hi[d] += np.float32(self.rlist)
            if any(o[d] == -1 for o in self.offsets):
                lo[d] -= np.float32(self.rlist)
        return lo, hi


def migrate_atoms_nd(t, box, grid, rank, rlist, home, x_home, to_tensor=None):
    """One repartitioning step of the N-D decomposition (dd_partition_system, domdec/partition.cpp; dd_redistribute_cg,
    domdec/redistribute.cpp), collective over the ranks of transport `t`: coordinates are wrapped into the box, atoms whose
    coordinates now lie in another domain go to its owner (at most one domain away per dimension, as in the reference), and
    every rank tells each rank that sees it at a half-shell offset which of its atoms lie near the shared boundary.
    Returns (home ascending, x_home, send_locals[k], recv_ids[k]) with k running over half_shell_offsets(grid) -- the pieces of
    DomainPlanND.from_parts, equal to what DomainPlanND computes from the global coordinates."""
    import torch
    from .domdec import wrap_into_box
    if to_tensor is None:
        to_tensor = torch.from_numpy
    geo = DomainPlanND.__new__(DomainPlanND)
    geo._geometry(box, grid, rank, rlist)
    home = np.ascontiguousarray(home, dtype=np.int32)
    x = wrap_into_box(x_home, geo.box)
    owner = DomainPlanND.owner_of(x, geo.box, geo.grid)
    stay = owner == rank
    # destination must be a neighbour domain: per dimension the same cell or the next one (periodically)
    for d in range(3):
        n = geo.grid[d]
        cd = np.array([DomainPlanND.coords_of(int(o), geo.grid)[d] for o in np.unique(owner)])
        dist = np.minimum((cd - geo.coords[d]) % n, (geo.coords[d] - cd) % n)
        if np.any(dist > 1):
            raise InputException("an atom moved more than one domain between two repartitioning steps")

    def pack(mask):
        m = np.nonzero(mask)[0]
        buf = np.empty((len(m), 4), np.int32)
        buf[:, 0] = home[m]
        buf[:, 1:] = x[m].view(np.int32)
        return buf

    dests = [int(r) for r in np.unique(owner[~stay])]
    out = {r: pack(owner == r) for r in dests}
    counts = t.allgather_object({r: len(b) for r, b in out.items()})
    srcs = [r for r in range(geo.nranks) if r != rank and counts[r].get(rank, 0) > 0]
    inb = {r: np.empty((counts[r][rank], 4), np.int32) for r in srcs}

    def exchange(send_list, recv_list):
        sends = [(to_tensor(np.ascontiguousarray(a)), dst) for a, dst in send_list if len(a)]
        recvs = [(to_tensor(a), src) for a, src in recv_list if len(a)]
        t.exchange(sends, recvs)
        k = 0
        for a, src in recv_list:
            if len(a):
                a[...] = recvs[k][0].cpu().numpy()
                k += 1

    exchange([(out[r], r) for r in sorted(out)], [(inb[r], r) for r in sorted(inb)])
    arrived = np.concatenate([inb[r] for r in sorted(inb)]) if inb else np.zeros((0, 4), np.int32)
    ids = np.concatenate([home[stay], arrived[:, 0]]).astype(np.int32)
    xs = np.concatenate([x[stay], arrived[:, 1:].copy().view(np.float32)]).astype(np.float32)
    order = np.argsort(ids, kind="stable")
    home_new, x_new = ids[order], np.ascontiguousarray(xs[order])
    if len(home_new) and not np.all(DomainPlanND.owner_of(x_new, geo.box, geo.grid) == rank):
        raise InputException("repartitioning left an atom outside its new owner's domain")
    # the new halo: per half-shell offset, my boundary atoms go to the rank that sees me there; sizes first
    send_locals = [geo.boundary_atoms(x_new, geo.coords, o).astype(np.int32) for o in geo.offsets]
    peers = [geo._peers(o) for o in geo.offsets]
    nsend = t.allgather_object([int(len(sl)) for sl in send_locals])
    recv_ids = [np.empty(nsend[nb][k], np.int32) for k, (nb, _, _, _) in enumerate(peers)]
    exchange([(home_new[sl], dst) for sl, (_, _, dst, _) in zip(send_locals, peers)],
             [(buf, nb) for buf, (nb, _, _, _) in zip(recv_ids, peers)])
    return home_new, x_new, send_locals, recv_ids


class DomainRankND:
    """One rank of the 1-D / 2-D / 3-D decomposed calculation: schedule of do_force() restricted to the nonbonded path,

        x home -> grid | local kernel | halo x from every half-shell neighbour | halo x -> grid | non-local kernel |
        halo f back to the owners, added (+ shift forces for images across a periodic edge) | forces -> atom order

    with the halo traffic on the transport (see the module docstring)."""

    def __init__(self, system, options, transport, grid, rank=None, device=0, use_windows=True):
        import torch
        if np.any(np.asarray(getattr(system, "box_offdiag", np.zeros(3))) != 0):
            raise InputException("domain decomposition of a triclinic cell is not supported")
        self.torch = torch
        self.t = transport
        self.rank = transport.rank if rank is None else rank
        self.options = options
        rc = float(options.pairlistCutoff)
        self.rlist = float(options.rlistOuter or rc)
        self.plan = p = DomainPlanND(system.x, system.box, grid, self.rank, self.rlist)
        self.nranks = p.nranks
        self.use_windows = bool(use_windows) and self.nranks > 1 and len(p.offsets) <= _lib.DD_MAX_LINKS
        self._windows_open = False
        self.nb = _lib.NbnxmGpu(device)
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.nb.set_stream(self.stream.cuda_stream)
        if hasattr(transport, "sync"):
            transport.sync = self.nb.synchronize
        configure_interactions(self.nb, system.nbfp, options, self.rlist)
        # the topology is global and replicated on every rank; ownership is what moves (repartition)
        self.topology = (np.asarray(system.types), np.asarray(system.q), np.asarray(system.excl_off), np.asarray(system.excl_idx))
        # periodic images only along dimensions that are not decomposed: across the others the images arrive as halo atoms
        self.nb.set_box(system.box, pbc=tuple(1 if g == 1 else 0 for g in p.grid))
        self.fshift_halo = np.zeros((_lib.SHIFTS, 3), np.float64)
        self._setup_local(np.ascontiguousarray(system.x[p.home]))

    def _setup_local(self, x_home, atoms_installed=False):
        """(re)build everything that depends on which atoms this rank owns and receives.  atoms_installed: the local topology is
        already in the context (built on the device by b200nb_dd_set_local_atoms) and x_home is a device tensor."""
        torch = self.torch
        p = self.plan
        if not atoms_installed:
            types, q, eo, ei = p.local_topology(*self.topology)
            self.nb.set_atoms(types, q, eo, ei)
        self.nlocal = p.nhome + p.nhalo
        box = p.box
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.x = torch.zeros((self.nlocal, 3), dtype=torch.float32, device=self.device)
            self.f = torch.zeros((self.nlocal, 3), dtype=torch.float32, device=self.device)
            if isinstance(x_home, torch.Tensor):
                self.x[:p.nhome].copy_(x_home)
            else:
                self.x[:p.nhome].copy_(torch.from_numpy(np.ascontiguousarray(x_home, dtype=np.float32)))
            self.send_idx = [torch.from_numpy(s["local"]).to(self.device) for s in p.send]
            self.send_buf = [torch.zeros((len(s["local"]), 3), dtype=torch.float32, device=self.device) for s in p.send]
            self.recv_f = [torch.zeros((len(s["local"]), 3), dtype=torch.float32, device=self.device) for s in p.send]
        self.send_shift = [(s["shift"].astype(np.float32) * box).astype(np.float32) for s in p.send]
        off = p.nhome
        self.recv_range = []
        for r in p.recv:
            self.recv_range.append((off, off + len(r["ids"])))
            off += len(r["ids"])
        self.nb.synchronize()
        if self.use_windows and not self._windows_open:
            self._open_windows()
        self.search()

    # -- peer-memory halo windows: one per rank, opened once by every neighbour (CUDA IPC across processes) ---------------------
    def _open_windows(self):
        import os
        p = self.plan
        nsend = sum(len(s["local"]) for s in p.send)
        self.max_halo, self.max_send = int(1.5 * p.nhalo) + 4096, int(1.5 * nsend) + 4096
        handle, ptr = self.nb.dd_create_window(self.max_halo, self.max_send)
        info = self.t.allgather_object(dict(pid=os.getpid(), handle=handle, ptr=ptr, max_halo=self.max_halo))
        # every rank I exchange with over some link, each opened once; peer number = position in this list
        self.peers = sorted({s["rank"] for s in p.send} | {r["rank"] for r in p.recv})
        if len(self.peers) > _lib.DD_MAX_PEERS:
            raise InputException("more than %d neighbour ranks" % _lib.DD_MAX_PEERS)
        for k, peer in enumerate(self.peers):
            pi = info[peer]
            if pi["pid"] == os.getpid():
                self.nb.dd_open_peer(k, window_ptr=pi["ptr"], peer_max_halo=pi["max_halo"])
            else:
                self.nb.dd_open_peer(k, ipc_handle=pi["handle"], peer_max_halo=pi["max_halo"])
        self._windows_open = True
        self.t.barrier()

    def _set_links(self):
        """The halo plan of this search interval as links (b200nb_dd_set_links): link k = half-shell offset k on every rank."""
        p = self.plan
        nsend = [len(s["local"]) for s in p.send]
        nrecv = [len(r["ids"]) for r in p.recv]
        if p.nhalo > self.max_halo or sum(nsend) > self.max_send:
            raise InputException("repartition: halo of %d / %d atoms exceeds the window capacity %d / %d"
                                 % (p.nhalo, sum(nsend), self.max_halo, self.max_send))
        counts = self.t.allgather_object(dict(nsend=nsend, nrecv=nrecv))
        links = []
        for k, (s, r) in enumerate(zip(p.send, p.recv)):
            dst, src = s["rank"], r["rank"]
            links.append(dict(send_peer=self.peers.index(dst), send_idx=s["local"], shift=self.send_shift[k],
                              # my atoms become halo atoms of `dst` after everything it receives over its links before k
                              peer_halo_offset=int(sum(counts[dst]["nrecv"][:k])),
                              recv_peer=self.peers.index(src), nrecv=nrecv[k],
                              # the forces on them return to the send entries of `src`, behind those of its links before k
                              peer_entry_offset=int(sum(counts[src]["nsend"][:k])),
                              # images across a periodic edge: the forces on them also enter that shift's shift force
                              # (domdec/domdec.cpp:426-458)
                              fshift_index=shift_index(r["shift"]) if np.any(r["shift"] != 0) else -1))
        self.nb.dd_set_links(p.nhome, p.nhalo, links)
        self.t.barrier()  # every rank has its plan before anybody steps

    def repartition(self, x_home=None, on_device=True):
        """Pair-search step with atom migration: the current coordinates of the home atoms (self.x[:nhome], on the device)
        decide the new owners; halo lists, grids and pair list are rebuilt.  Collective.  Returns the new plan.
        on_device: the decisions are taken by the kernels of csrc/dd_partition.cu (coordinates stay on the GPU); False: the numpy
        restatement of the same step (migrate_atoms_nd), what the gloo tests run and the kernels are checked against."""
        torch = self.torch
        p = self.plan
        self.nb.synchronize()
        if on_device:
            return self._repartition_device(x_home)
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
        home, x_new, send_locals, recv_ids = migrate_atoms_nd(self.t, p.box, p.grid, self.rank, self.rlist, p.home, x_home,
                                                              to_tensor=to_dev)
        self.plan = DomainPlanND.from_parts(p.box, p.grid, self.rank, self.rlist, home, send_locals, recv_ids)
        self._setup_local(x_new)
        return self.plan

    def _repartition_device(self, x_home):
        """repartition() with the coordinate-dependent work on the GPU: wrap + new owner cell per atom as one of the 27 neighbour
        offsets (k_wrap_classify_nd), stable 28-way partition into stayers and leavers per offset, messages per destination rank,
        merge into the new home set, per half-shell offset the boundary atoms the neighbour there needs (k_select_boundary +
        partition), local topology through the device-resident global -> local look-up.  The host sees counts and the index
        lists of the plan (global indices of home / halo atoms, send lists): bookkeeping, no coordinates."""
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
            gid = torch.from_numpy(np.ascontiguousarray(p.home, dtype=np.int32)).to(dev)
            n = p.nhome
            code = torch.empty(max(n, 1), **i32)
            idx = torch.empty(max(n, 1), **i32)
            nb.dd_wrap_classify_nd(x.data_ptr(), n, p.box, p.grid, p.coords, code.data_ptr())
            cnt = nb.dd_partition_indices(code.data_ptr(), n, 28, idx.data_ptr())
            if cnt[27]:
                raise InputException("an atom moved more than one domain between two repartitioning steps")
            start = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
            nstay = cnt[13]
            # messages per destination RANK: the offsets that lead to the same rank (two cells along a dimension) are concatenated
            out = {}
            for c in range(27):
                if c == 13 or cnt[c] == 0:
                    continue
                o = (c // 9 - 1, (c // 3) % 3 - 1, c % 3 - 1)
                dst, _ = p.neighbour(p.coords, o, +1)
                buf = torch.empty((cnt[c], 4), **i32)
                nb.dd_pack_atoms(idx.data_ptr() + 4 * int(start[c]), cnt[c], gid.data_ptr(), x.data_ptr(), buf.data_ptr())
                out.setdefault(int(dst), []).append(buf)
            out = {r: (torch.cat(b).contiguous() if len(b) > 1 else b[0]) for r, b in out.items()}
            counts = t.allgather_object({r: int(b.shape[0]) for r, b in out.items()})
            srcs = [r for r in range(p.nranks) if r != self.rank and counts[r].get(self.rank, 0) > 0]
            inb = {r: torch.empty((counts[r][self.rank], 4), **i32) for r in srcs}
            t.exchange([(out[r], r) for r in sorted(out)], [(inb[r], r) for r in sorted(inb)])
            arrived = torch.cat([inb[r] for r in sorted(inb)]).contiguous() if inb else torch.empty((0, 4), **i32)
            m = int(arrived.shape[0])
            n_new = nstay + m
            gid_new = torch.empty(max(n_new, 1), **i32)[:n_new]
            x_new = torch.empty((max(n_new, 1), 3), dtype=torch.float32, device=dev)[:n_new]
            nb.dd_merge_home(idx.data_ptr() + 4 * int(start[13]), nstay, gid.data_ptr(), x.data_ptr(), arrived.data_ptr() if m else 0, m,
                             gid_new.data_ptr(), x_new.data_ptr())
            code2 = torch.empty(max(n_new, 1), **i32)
            idx2 = torch.empty(max(n_new, 1), **i32)
            nb.dd_wrap_classify_nd(x_new.data_ptr(), n_new, p.box, p.grid, p.coords, code2.data_ptr())
            if nb.dd_partition_indices(code2.data_ptr(), n_new, 28, idx2.data_ptr())[13] != n_new:
                raise InputException("repartitioning left an atom outside its new owner's domain")
            # the new halo: per half-shell offset, my boundary atoms go to the rank that sees me there; sizes first
            send_locals, send_gids = [], []
            for o in p.offsets:
                nb.dd_select_boundary(x_new.data_ptr(), n_new, p.lo, p.hi, o, self.rlist, code2.data_ptr())
                nkeep, nsend = nb.dd_partition_indices(code2.data_ptr(), n_new, 2, idx2.data_ptr())
                sl = idx2[nkeep:nkeep + nsend].clone()
                sg = torch.empty(max(nsend, 1), **i32)[:nsend]
                nb.dd_gather_int(sl.data_ptr() if nsend else 0, nsend, gid_new.data_ptr(), sg.data_ptr() if nsend else 0)
                send_locals.append(sl)
                send_gids.append(sg)
            peers = [p._peers(o) for o in p.offsets]
            nsends = t.allgather_object([int(sl.shape[0]) for sl in send_locals])
            recv_ids = [torch.empty(max(nsends[nbr][k], 1), **i32)[:nsends[nbr][k]] for k, (nbr, _, _, _) in enumerate(peers)]
            t.exchange([(sg, dst) for sg, (_, _, dst, _) in zip(send_gids, peers)], [(buf, nbr) for buf, (nbr, _, _, _) in zip(recv_ids, peers)])
            local_gid = torch.cat([gid_new] + recv_ids).contiguous()
            nb.dd_set_local_atoms(local_gid.data_ptr(), int(local_gid.shape[0]))
            nb.synchronize()
            home_h = gid_new.cpu().numpy()
            send_h = [sl.cpu().numpy() for sl in send_locals]
            recv_h = [r.cpu().numpy() for r in recv_ids]
        self.plan = DomainPlanND.from_parts(p.box, p.grid, self.rank, self.rlist, home_h, send_h, recv_h)
        self._setup_local(x_new, atoms_installed=True)
        return self.plan

    def search(self):
        p = self.plan
        self._halo_x()
        self.nb.synchronize()
        lower = np.where(np.array(p.grid) > 1, p.lo, 0.0).astype(np.float32)
        upper = np.where(np.array(p.grid) > 1, p.hi, p.box).astype(np.float32)
        self.nb.put_on_grid(self.x.data_ptr(), lower, upper, 0, 0, p.nhome, on_device=True)
        if p.nhalo:
            hl, hu = p.halo_bounds()
            hl = np.where(np.array(p.grid) > 1, hl, 0.0).astype(np.float32)
            hu = np.where(np.array(p.grid) > 1, hu, p.box).astype(np.float32)
            self.nb.put_on_grid(self.x.data_ptr(), hl, hu, 1, p.nhome, self.nlocal, on_device=True)
        self.nb.build_pairlist()
        if self.use_windows:
            self._set_links()

    def _exchange(self, sends, recvs):
        """all messages of one halo phase at once; messages between the same two ranks keep the half-shell offset order on
        both sides, so they match without tags"""
        if hasattr(self.t, "exchange"):
            self.t.exchange(sends, recvs)
        else:
            raise InputException("the transport has no exchange(): use domdec.TorchDistTransport or LoopbackTransport")

    def _halo_x(self):
        p = self.plan
        if not p.recv:
            return
        torch = self.torch
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            sends, recvs = [], []
            for k, s in enumerate(p.send):
                n = len(s["local"])
                if n:
                    self.nb.halo_pack_x(self.x.data_ptr(), self.send_idx[k].data_ptr(), n, self.send_shift[k], self.send_buf[k].data_ptr())
                sends.append((self.send_buf[k], s["rank"]))
            for k, r in enumerate(p.recv):
                a, b = self.recv_range[k]
                recvs.append((self.x[a:b], r["rank"]))
            self._exchange(sends, recvs)

    def _halo_f(self, virial):
        p = self.plan
        if not p.recv:
            return
        torch = self.torch
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            sends, recvs = [], []
            for k, r in enumerate(p.recv):
                a, b = self.recv_range[k]
                sends.append((self.f[a:b], r["rank"]))
            for k, s in enumerate(p.send):
                recvs.append((self.recv_f[k], s["rank"]))
            self._exchange(sends, recvs)
            self.fshift_halo[:] = 0
            for k, s in enumerate(p.send):
                n = len(s["local"])
                if not n:
                    continue
                self.nb.halo_unpack_f(self.f.data_ptr(), self.send_idx[k].data_ptr(), n, self.recv_f[k].data_ptr())
                if virial and np.any(s["shift"] != 0):
                    # domdec/domdec.cpp:426-458: forces on images that crossed a periodic edge also enter the shift forces
                    self.fshift_halo[shift_index(s["shift"])] += self.recv_f[k].sum(0, dtype=torch.float64).cpu().numpy()

    def step(self, flags=0):
        """One nonbonded step on the coordinates in self.x[:nhome] (device); leaves forces in self.f[:nhome]."""
        p = self.plan
        if self.use_windows:
            self.nb.dd_step(self.x.data_ptr(), self.f.data_ptr(), flags)
            return
        self.nb.set_x(self.x.data_ptr(), on_device=True, atom_begin=0, atom_end=p.nhome)
        self.nb.clear_outputs()
        self.nb.launch_force(0, flags)
        self._halo_x()
        if p.nhalo:
            self.nb.set_x(self.x.data_ptr(), on_device=True, atom_begin=p.nhome, atom_end=self.nlocal)
            self.nb.launch_force(1, flags)
        self.nb.get_f(self.f.data_ptr(), on_device=True)
        self._halo_f(bool(flags & _lib.FLAG_VIRIAL))

    def compute(self, x_home_host, flags=0, f_home_host=None):
        """Host coordinates of the home atoms in, (f_home, fshift[45,3], e_lj, e_el) out: this rank's share of the sums."""
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
            if not self.use_windows:
                fs = fs + self.fshift_halo.astype(np.float32)
        return fh, fs, elj, eel

    def pair_count(self, r):
        return self.nb.pair_count(r)

    def close(self):
        self.nb.synchronize()
        self.nb.close()
        self.x = self.f = None
        self.send_idx = self.send_buf = self.recv_f = None
