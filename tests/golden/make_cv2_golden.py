"""Known answers for Interfaces / Propensities from the UNMODIFIED reference (build container only).

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_cv2_golden.py -> tests/golden/cv2_known_answers.npz
"""
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")

import networkx as nx  # noqa: E402
import numpy as np  # noqa: E402

from sponet import CNVMParameters, Interfaces  # noqa: E402
from sponet.collective_variables import Propensities  # noqa: E402
from sponet.utils import calculate_neighbor_list  # noqa: E402

g = nx.barabasi_albert_graph(120, 3, seed=4)
nl = calculate_neighbor_list(g)
deg = np.array([len(nb) for nb in nl])
row_ptr = np.zeros(121, dtype=np.int32)
np.cumsum(deg, out=row_ptr[1:])
col = np.concatenate(nl).astype(np.int32)
x = np.random.default_rng(0).integers(0, 2, (9, 120))
r = np.array([[0, 1.3], [0.7, 0]])
rt = np.array([[0, 0.02], [0.05, 0]])
out = dict(row_ptr=row_ptr, col=col, x=x, r=r, r_tilde=rt, num_edges=g.number_of_edges(),
           interfaces=Interfaces(g)(x), interfaces_norm=Interfaces(g, True)(x), interfaces_1d=Interfaces(g)(x[0]))
for alpha in (1.0, 0.5, 0.0):
    p = CNVMParameters(num_opinions=2, network=nl, r=r, r_tilde=rt, alpha=alpha)
    out[f"prop_net_a{alpha}"] = Propensities(p)(x)
    out[f"prop_net_norm_a{alpha}"] = Propensities(p, True)(x)
    pc = CNVMParameters(num_opinions=2, num_agents=120, r=r, r_tilde=rt, alpha=alpha)
    out[f"prop_complete_a{alpha}"] = Propensities(pc)(x)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv2_known_answers.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
