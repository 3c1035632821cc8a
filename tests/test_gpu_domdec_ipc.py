"""The decomposed step across PROCESSES: two / three ranks, each its own process on GPU 0, halo windows mapped through CUDA IPC
(cudaIpcGetMemHandle / cudaIpcOpenMemHandle) -- the configuration bench.py --gpus N runs, minus the second GPU.  Forces, energies
and pair count against the single-domain oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import gmxapi_b200 as g
from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RC = 0.9
ENERGY_TOL = 2e-5  # relative, against the oracle's double-precision sums (the reference's own FP32 sums sit 1e-5 .. 2e-4 from those)


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return str(s.getsockname()[1])


@pytest.mark.parametrize("nranks", [2, 3])
def test_decomposed_step_over_cuda_ipc_windows(built, tmp_path, nranks):
    workload, port = "water_24k", free_port()
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", B200NB_DD_PUSH_INLINE="0")  # two processes: the product path
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dd_ipc_worker.py"), str(r), str(nranks), port, workload,
                               str(tmp_path / ("rank%d.npz" % r))], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(nranks)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    res = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(nranks)]
    assert len({int(r["pid"]) for r in res}) == nranks  # really separate processes: the windows went through IPC handles
    s = g.systems.named(workload)
    beta = float(np.float32(g.systems.ewald_beta(RC)))
    fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, RC, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD, beta=beta)
    f = np.zeros((s.n, 3), np.float64)
    seen = np.zeros(s.n, int)
    for r in res:
        assert int(r["nhalo"]) > 0
        f[r["home"]] = r["f"]
        seen[r["home"]] += 1
    assert np.all(seen == 1)
    assert np.sqrt(((f - fo) ** 2).sum() / (fo ** 2).sum()) < 1e-5
    # pairs that cross the periodic x edge may flip by the rounding of the shifted coordinates (tests/test_gpu_domdec.py)
    assert abs(sum(int(r["npairs"]) for r in res) - npairs) <= 2
    assert abs(sum(float(r["elj"]) for r in res) - evo) <= ENERGY_TOL * abs(evo)
    assert abs(sum(float(r["eel"]) for r in res) - eco) <= ENERGY_TOL * abs(eco)
