"""Steady-state reference values for tests/test_gpu_sweep.py (build container only).

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_steady_golden.py

The UNMODIFIED reference's `estimate_steady_state` (sponet/steady_state.py:12-81; its generator is a module-level
`default_rng()` and cannot be seeded) is run 8 times per configuration with tol_abs = 0.005; the fixture stores the
individual estimates, their mean, and the inputs.  The GPU test requires its own estimate to lie within the
reference's tolerance of that mean."""
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")

import networkx as nx  # noqa: E402
import numpy as np  # noqa: E402

from sponet import CNTMParameters, CNVMParameters, OpinionShares, estimate_steady_state  # noqa: E402
from sponet.utils import calculate_neighbor_list  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)


def csr(nl):
    deg = np.array([len(nb) for nb in nl])
    rp = np.zeros(len(nl) + 1, dtype=np.int32)
    np.cumsum(deg, out=rp[1:])
    return rp, np.concatenate([np.asarray(nb, dtype=np.int32) for nb in nl])


g = nx.fast_gnp_random_graph(400, 8 / 399, seed=11)
for v in list(nx.isolates(g)):
    g.add_edge(v, (v + 1) % 400)
nl = calculate_neighbor_list(g)
rp, col = csr(nl)
p = CNVMParameters(num_opinions=3, network=nl, r=R3, r_tilde=RT3)
est = np.array([estimate_steady_state(p, OpinionShares(3, normalize=True), tol_abs=0.005) for _ in range(8)])
np.savez_compressed(os.path.join(OUT, "steady_cnvm_er400.npz"), model="cnvm", row_ptr=rp, col=col, r=R3, r_tilde=RT3,
                    num_opinions=3, estimates=est, mean=est.mean(0), tol_abs=0.005)
print("cnvm", est.mean(0), est.std(0))

pt = CNTMParameters(network=g, r=1.0, r_tilde=0.3, threshold_01=0.4, threshold_10=0.6)
est = np.array([estimate_steady_state(pt, OpinionShares(2, normalize=True), tol_abs=0.005) for _ in range(8)])
np.savez_compressed(os.path.join(OUT, "steady_cntm_er400.npz"), model="cntm", row_ptr=rp, col=col, r=1.0, r_tilde=0.3,
                    threshold_01=0.4, threshold_10=0.6, num_opinions=2, estimates=est, mean=est.mean(0), tol_abs=0.005)
print("cntm", est.mean(0), est.std(0))
