"""Known answers for the host-side logic, produced by the UNMODIFIED reference (build container only).

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_host_golden.py  ->  tests/golden/host_logic.json
"""
import json
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

from sponet import CNVMParameters  # noqa: E402
from sponet.multiprocessing import _split_runs  # noqa: E402
from sponet.sampling import build_alias_table  # noqa: E402
from sponet.utils import argmatch, counts_from_shares, mask_subsequent_duplicates, t_eval_to_ndarray  # noqa: E402

rng = np.random.default_rng(2026)
out = {"rates_style1": [], "rates_style2": [], "rate_errors": [], "argmatch": [], "counts_from_shares": [],
       "split_runs": [], "mask_dups": [], "t_eval": [], "t_eval_errors": [], "alias": []}


def tolist(a):
    return np.asarray(a, dtype=float).tolist()


style1 = [
    (3, 1, 0.1), (None, [[1, 1], [2, 3]], [[0.1, 0], [0.3, -2]]), (None, [[0, 1], [2, 0]], 0.1),
    (None, 2, [[0, 0.1], [0.2, 0]]), (3, [[0, 1, 2], [1, 0, 1], [2, 0, 0]], [[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]]),
    (2, 0, 0.3), (2, 1.5, 0), (4, rng.random((4, 4)).tolist(), (0.1 * rng.random((4, 4))).tolist()),
]
for M, r, rt in style1:
    p = CNVMParameters(M, r=r, r_tilde=rt, num_agents=10)
    out["rates_style1"].append({"num_opinions": M, "r_in": r, "r_tilde_in": rt, "M": p.num_opinions, "r": tolist(p.r),
                                "r_tilde": tolist(p.r_tilde), "r_imit": float(p.r_imit), "r_noise": float(p.r_noise),
                                "prob_imit": tolist(p.prob_imit), "prob_noise": tolist(p.prob_noise)})
style2 = [(3, 2.0, 0.6, 1, 1), (None, 1.0, 0.2, [[0, 0.5], [1, 0]], [[0, 1], [0.25, 0]]), (2, 0.0, 1.0, 0.5, 1)]
for M, ri, rn, pi, pn in style2:
    p = CNVMParameters(M, num_agents=10, r_imit=ri, r_noise=rn, prob_imit=pi, prob_noise=pn)
    out["rates_style2"].append({"num_opinions": M, "r_imit": ri, "r_noise": rn, "prob_imit_in": pi, "prob_noise_in": pn,
                                "M": p.num_opinions, "r": tolist(p.r), "r_tilde": tolist(p.r_tilde),
                                "prob_imit": tolist(p.prob_imit), "prob_noise": tolist(p.prob_noise)})
errors = [
    dict(num_opinions=3, r=1, r_tilde=1),  # no network
    dict(num_agents=10, r=1, r_tilde=1),  # num_opinions missing
    dict(num_agents=10, num_opinions=2, r=-1, r_tilde=1),
    dict(num_agents=10, num_opinions=3, r=[[0, 1], [1, 0]], r_tilde=0.1),
    dict(num_agents=10, r=[[0, 1, 2], [1, 0, 1]], r_tilde=0.1),
    dict(num_agents=10, r=[[0, 1], [1, 0]], r_tilde=[[0, 1, 1], [1, 0, 1], [1, 1, 0]]),
    dict(num_agents=10, num_opinions=2),  # no rates
    dict(num_agents=10, num_opinions=2, r_imit=1, r_noise=1, prob_imit=1.5),
    dict(num_agents=10, num_opinions=2, r_imit=-1, r_noise=1),
]
for kw in errors:
    try:
        CNVMParameters(**kw)
        raised = None
    except Exception as e:  # noqa: BLE001
        raised = type(e).__name__
    out["rate_errors"].append({"kwargs": kw, "raises": raised})

for t_ref, t in [([1.8, 2, 4.4, 6.7, 6.9], [1, 2, 3, 4, 5, 6, 7]), ([0.1, 1.8, 2, 2.5, 4.4, 6.7, 7.7], [1, 2, 3, 4, 5]),
                 (np.sort(rng.random(20) * 10).tolist(), np.sort(rng.random(9) * 10).tolist())]:
    t_ref, t = np.array(t_ref, dtype=float), np.array(t, dtype=float)
    if t_ref[0] < t[0]:
        pass
    out["argmatch"].append({"x_ref": t_ref.tolist(), "x": t.tolist(), "out": argmatch(t_ref, t).tolist()})
for shares, n in [([[0.2, 0.3, 0.5], [0.209, 0.309, 0.482]], 100), ([0.204, 0.304, 0.492], 100), ([0.555, 0.445, 0], 100),
                  ([0.554, 0.446, 0], 10), ([0.155, 0.155, 0.155, 0.155, 0.380], 100), ([0.055] * 18 + [0.010], 100)]:
    out["counts_from_shares"].append({"shares": shares, "num_agents": n, "counts": counts_from_shares(shares, n).tolist()})
for a, b in [(20, 6), (5, 8), (100, 1), (0, 3), (7, 7)]:
    out["split_runs"].append({"num_runs": a, "num_chunks": b, "chunks": _split_runs(a, b).tolist()})
for x in [[1, 1, 2, 2, 2, 3, 1, 3, 3], [[1, 1], [1, 1], [2, 2], [2, 1], [2, 1], [1, 1]]]:
    out["mask_dups"].append({"x": x, "mask": mask_subsequent_duplicates(np.array(x)).tolist()})
for te, tm in [(11, 10.0), (2, 0.3), ([0.5, 1.0, 3.0], 3.0), ([0, 1], 5)]:
    out["t_eval"].append({"t_eval": te, "t_max": tm, "out": t_eval_to_ndarray(te, tm).tolist()})
for te, tm in [(1.5, 10.0), ([], 1.0), ([1, 1], 2.0), ([2, 1], 2.0), ([-1, 1], 2.0), ([0, 3], 2.0)]:
    try:
        t_eval_to_ndarray(te, tm)
        raised = None
    except Exception as e:  # noqa: BLE001
        raised = [type(e).__name__, str(e)]
    out["t_eval_errors"].append({"t_eval": te, "t_max": tm, "raises": raised})
for w in [rng.random(50) ** 3, np.ones(7), np.array([1.0, 2.0, 3.0, 4.0]), rng.integers(1, 30, 200).astype(float) ** 0.5]:
    prob, alias = build_alias_table(np.asarray(w, dtype=float))
    out["alias"].append({"weights": [x.hex() for x in np.asarray(w, dtype=float)], "prob": [float(x).hex() for x in prob],
                         "alias": alias.tolist()})

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_logic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print("wrote host_logic.json", {k: len(v) for k, v in out.items()})
