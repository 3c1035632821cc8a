"""Worker of tests/test_gpu_domdec_ipc.py: ONE rank of a decomposed run in its OWN process (rank, nranks, port, workload, out path
on the command line).  All ranks share GPU 0, so the halo windows of the neighbours are reached through CUDA IPC mappings --
the path the multi-GPU benchmark runs, which the thread-based tests (one process, plain device pointers) do not exercise.
Set-up data travels over gloo; search-step halo coordinates are staged through the host (gloo has no CUDA point-to-point)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, nranks, port, workload, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(nranks))
    import torch
    import torch.distributed as dist
    import gmxapi_b200 as g
    from gmxapi_b200 import lib as nb
    from gmxapi_b200.domdec import DomainRank, TorchDistTransport
    dist.init_process_group("gloo", rank=rank, world_size=nranks)

    class HostStaged(TorchDistTransport):
        def sendrecv(self, send, dst, recv, src):
            s = send.cpu() if send is not None and send.numel() else None
            r = torch.empty(recv.shape, dtype=recv.dtype) if recv is not None and recv.numel() else None
            super().sendrecv(s, dst, r, src)
            if r is not None:
                recv.copy_(r)

    s = g.systems.named(workload)
    opt = g.NBKernelOptions(pairlistCutoff=0.9, coulombType=g.CoulombType.Pme, computeVirialAndEnergy=True)
    d = DomainRank(s, opt, HostStaged(), device=0, use_windows=True)
    assert d.use_windows
    x_pin = torch.from_numpy(np.ascontiguousarray(s.x[d.plan.home])).pin_memory()
    for _ in range(4):  # repeated steps: flags advance, windows are overwritten
        f, fs, elj, eel = d.compute(x_pin, nb.FLAG_ENERGY | nb.FLAG_VIRIAL)
    n = d.pair_count(0.9)
    np.savez(out, home=d.plan.home, f=f.numpy(), elj=elj, eel=eel, npairs=n, nhalo=d.plan.nhalo, pid=os.getpid())
    dist.barrier()
    d.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
