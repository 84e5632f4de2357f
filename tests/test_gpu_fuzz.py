"""GPU: randomized differential test of the native kernels against the sequential oracle.

A hundred and ten seeded configurations drawn at random -- graph family and size, number of opinions (all four packing
widths), alpha (uniform / alias agent selection), irregular output grids with empty intervals, warp schedule,
shared- or global-memory chain state -- each compared bit for bit (states, CV values, event counters).
"""
import os

import numpy as np
import pytest

from conftest import library_neighbor_table

pytestmark = pytest.mark.gpu

# SPONET_FUZZ_OFFSET=k shifts every seed: more configurations without editing the file
OFFSET = 100_000 * int(os.environ.get("SPONET_FUZZ_OFFSET", "0"))


def _graph(rng, n):
    import networkx as nx

    kind = rng.choice(["er", "ba", "ws", "cycle", "star"])
    s = int(rng.integers(1 << 30))
    if kind == "er":
        g = nx.fast_gnp_random_graph(n, min(1.0, float(rng.uniform(2, 12)) / (n - 1)), seed=s)
    elif kind == "ba":
        g = nx.barabasi_albert_graph(n, int(rng.integers(1, 6)), seed=s)
    elif kind == "ws":
        g = nx.watts_strogatz_graph(n, 4, 0.2, seed=s)
    elif kind == "cycle":
        g = nx.cycle_graph(n)
    else:
        g = nx.star_graph(n - 1)  # one hub of degree n - 1: no neighbour table beyond 64, every step has hazards
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % n)
    return g


def _t_eval(rng, mean_gap):
    """Irregular output grid: intervals expected to hold 0, a fraction of one, a few or a few thousand events
    (``mean_gap`` = mean time between two events of the chain, the reference's ``next_event_rate``)."""
    T = int(rng.integers(1, 9))
    events = rng.choice([0.0, 0.01, 5.0, 32.0, 33.0, 200.0, 3000.0], size=max(T - 1, 0)) * rng.uniform(0.5, 1.5, size=max(T - 1, 0))
    return np.concatenate(([0.0], np.cumsum(events * mean_gap)))


@pytest.mark.parametrize("case", range(80))
def test_cnvm_native_fuzz(oracle, case):
    from sponet_b200 import CNVM, CNVMParameters, engine

    rng = np.random.default_rng(1000 + case + OFFSET)
    n = int(rng.choice([33, 64, 100, 257, 1000, 4096, 5000]))
    M = int(rng.choice([2, 3, 4, 5, 16, 17, 60]))
    alpha = float(rng.choice([1.0, 1.0, 0.5, 0.0]))
    r = rng.random((M, M)) * rng.choice([0.0, 1.0], size=(M, M), p=[0.2, 0.8])
    rt = 0.1 * rng.random((M, M)) * rng.choice([0.0, 1.0], size=(M, M), p=[0.3, 0.7])
    if not (r + rt).any():
        r[0, 1] = 1.0
    complete = rng.random() < 0.15
    if complete:
        params = CNVMParameters(num_opinions=M, num_agents=n, r=r, r_tilde=rt, alpha=alpha)
        rp = col = None
    else:
        params = CNVMParameters(num_opinions=M, network=_graph(rng, n), r=r, r_tilde=rt, alpha=alpha)
        rp, col = oracle.neighbor_list_to_csr(params.network)
    model = CNVM(params)
    rate, noise_p = model.event_rates()
    t_eval = _t_eval(rng, rate)
    S, R = int(rng.integers(1, 3)), int(rng.integers(1, 5))
    states = rng.integers(0, M, (S, n)).astype(np.uint8)
    seed = int(rng.integers(1 << 62))
    wpc = int(rng.choice([0, 1, 2, 3, 5, 8]))
    force_global = bool(rng.random() < 0.25)
    struct, keep = model._struct()
    spec = (np.arange(M, dtype=np.int32), 1.0)
    counters = np.zeros(4, dtype=np.uint64)
    out_x, out_cv = engine.run_native("cnvm", model._graph(), n, struct, states, t_eval, seed, R, 0, R, cv_spec=spec,
                                      want_x=True, want_cv=True, counters=counters, force_global_state=force_global,
                                      warps_per_chain=wpc)
    thr_i, thr_n = oracle.prob_to_thr(params.prob_imit), oracle.prob_to_thr(params.prob_noise)
    noise_thr = int(oracle.prob_to_thr(noise_p))
    da = np.asarray(model.degree_alpha, dtype=float)
    alias_thr = alias = None
    if not complete and not np.all(da == da[0]):
        prob, alias = oracle.build_alias_table(da)
        alias_thr = oracle.prob_to_thr(prob)
    nbtab, nblg = None, 0
    if not complete:
        tab, lg = library_neighbor_table(model._graph())
        exp, exp_lg = oracle.build_neighbor_table(rp, col)
        assert lg == exp_lg
        if tab is not None:
            np.testing.assert_array_equal(tab, exp)
            nbtab, nblg = tab, lg
    ev_total = acc_total = 0
    for j in range(S):
        for i in range(R):
            chain = j * R + i
            counts = oracle.native_event_counts(seed, chain, 1, t_eval, rate)[0] if t_eval.size > 1 else np.zeros(0, dtype=np.int64)
            ox, oc, ev, acc = oracle.cnvm_native(states[j], M, rp, col, noise_thr, thr_i, thr_n, alias_thr, alias, counts,
                                                 seed, chain, nbtab=nbtab, nbtab_log2=nblg)
            np.testing.assert_array_equal(out_x[j, i], ox, err_msg=f"case {case}: n={n} M={M} alpha={alpha} wpc={wpc} global={force_global}")
            np.testing.assert_array_equal(out_cv[j, i], oc.astype(float))
            ev_total += ev
            acc_total += acc
    assert int(counters[0]) == ev_total and int(counters[1]) == acc_total
    del keep


@pytest.mark.parametrize("case", range(30))
def test_cntm_native_fuzz(oracle, case):
    from sponet_b200 import CNTM, CNTMParameters, engine

    rng = np.random.default_rng(5000 + case + OFFSET)
    n = int(rng.choice([33, 100, 1000, 5000]))
    thr01, thr10 = float(rng.choice([0.0, 0.25, 0.5, 0.75, 1.0, 1.5])), float(rng.choice([0.0, 0.3, 0.5, 1.0]))
    params = CNTMParameters(network=_graph(rng, n), r=float(rng.uniform(0.2, 2)), r_tilde=float(rng.choice([0.0, 0.1, 1.0])),
                            threshold_01=thr01, threshold_10=thr10)
    model = CNTM(params)
    rp, col = oracle.neighbor_list_to_csr(model.neighbor_list)
    rate = model.next_event_rate
    t_eval = _t_eval(rng, rate)
    S, R = int(rng.integers(1, 3)), int(rng.integers(1, 5))
    states = rng.integers(0, 2, (S, n)).astype(np.uint8)
    seed = int(rng.integers(1 << 62))
    counters = np.zeros(4, dtype=np.uint64)
    out_x, out_cv = engine.run_native("cntm", model._graph(), n, model._struct(), states, t_eval, seed, R, 0, R,
                                      cv_spec=(np.arange(2, dtype=np.int32), 1.0), want_x=True, want_cv=True,
                                      counters=counters, force_global_state=bool(rng.random() < 0.25))
    noise_thr = int(oracle.prob_to_thr(model.noise_prob))
    ev_total = acc_total = 0
    for j in range(S):
        for i in range(R):
            chain = j * R + i
            counts = oracle.native_event_counts(seed, chain, 1, t_eval, rate)[0] if t_eval.size > 1 else np.zeros(0, dtype=np.int64)
            ox, oc, ev, acc = oracle.cntm_native(states[j], rp, col, noise_thr, thr01, thr10, counts, seed, chain)
            np.testing.assert_array_equal(out_x[j, i], ox, err_msg=f"case {case}: n={n}")
            np.testing.assert_array_equal(out_cv[j, i], oc.astype(float))
            ev_total += ev
            acc_total += acc
    assert int(counters[0]) == ev_total and int(counters[1]) == acc_total
