import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def raw_stream(bitgen: str, seed: int, n: int) -> np.ndarray:
    """What the reference consumed: BitGenerator(seed).random_raw."""
    from numpy.random import PCG64, Philox

    return {"PCG64": PCG64, "Philox": Philox}[bitgen](int(seed)).random_raw(n)


def csr_to_neighbor_list(row_ptr, col):
    return [np.ascontiguousarray(col[row_ptr[i] : row_ptr[i + 1]], dtype=np.int32) for i in range(len(row_ptr) - 1)]


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc

    orc.build()
    return orc


def library_neighbor_table(graph):
    """(table uint64 [n << lg], lg) the library built for `graph` (an engine.GraphHandle), or (None, -1)."""
    import ctypes as C

    from sponet_b200 import _lib

    lg = C.c_int32(0)
    _lib.check(_lib.lib().sponet_graph_neighbor_table(graph.handle, C.byref(lg), None))
    if lg.value < 0:
        return None, -1
    tab = np.empty(graph.num_agents << lg.value, dtype=np.uint64)
    _lib.check(_lib.lib().sponet_graph_neighbor_table(graph.handle, C.byref(lg), tab.ctypes.data))
    return tab, lg.value
