import os
import sys

import pytest

# The single-GPU domain-decomposition tests run up to 6 ranks x 2 streams in ONE process; with CUDA's default of 8 hardware
# queues, streams alias and a flag-wait kernel could sit in front of the push kernel it waits for.  Production runs one
# process per GPU (2 streams).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Same cause, second measure: ranks that are THREADS of one process put the branches of all their step graphs on the hardware queues
# of one device, where a halo push on a graph branch of its own (the product path) can end up queued behind another rank's flag
# wait -- a dependency cycle that one process per GPU cannot have.  The in-process tests therefore push from the main stream;
# the separate branch is what tests/test_gpu_domdec_ipc.py (two processes) and every multi-GPU bench.py run (forces checked
# against a single-domain run) exercise.
os.environ.setdefault("B200NB_DD_PUSH_INLINE", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (cheap no-op when up to date)."""
    import __graft_entry__ as ge
    lib = os.path.join(ROOT, "gmxapi_b200", "libb200nb.so")
    orc = os.path.join(ROOT, "oracle", "_build", "libnbnxm_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        ge.build()
    return True
