"""GPU: the native-mode kernels against their sequential statement (oracle/oracle_native.inc).

The kernels evaluate 32 candidate events per warp step in parallel; the oracle executes the same
Philox-keyed chain one event at a time.  Results must be identical (integer states / counts).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import library_neighbor_table

pytestmark = pytest.mark.gpu


def _nbtab(oracle, model, rp, col):
    """The neighbour sampler table the library uses for this model's graph, cross-checked against the
    oracle's independent construction.  (None, 0) for complete graphs / graphs without a table."""
    if rp is None:
        return None, 0
    tab, lg = library_neighbor_table(model._graph())
    exp, exp_lg = oracle.build_neighbor_table(rp, col)
    assert lg == exp_lg
    if tab is None:
        return None, 0
    np.testing.assert_array_equal(tab, exp)
    return tab, lg

R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)


def _net(kind, n, seed):
    import networkx as nx

    if kind == "er":
        g = nx.erdos_renyi_graph(n, 10 / (n - 1), seed=seed)
    elif kind == "ba":
        g = nx.barabasi_albert_graph(n, 5, seed=seed)
    else:
        g = nx.cycle_graph(n)
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % n)
    return g


def test_event_counts_match_oracle(oracle):
    from sponet_b200 import _lib

    t_eval = np.array([0.0, 0.1, 0.25, 0.2500001, 1.0, 1.00002, 3.0])
    rate = 1.0 / 2.6e4  # lam from 2.6e-3 to 5.2e4 across the intervals
    out = np.empty((40, t_eval.size - 1), dtype=np.int64)
    _lib.check(_lib.lib().sponet_native_event_counts(99, 1000, 40, _lib.ptr(t_eval), t_eval.size, rate, _lib.ptr(out), None))
    np.testing.assert_array_equal(out, oracle.native_event_counts(99, 1000, 40, t_eval, rate))


@pytest.mark.parametrize("kind,n,M,alpha,force_global", [
    ("er", 1000, 3, 1.0, False),
    ("er", 1000, 3, 1.0, True),
    ("ba", 600, 3, 0.5, False),   # alias tables
    ("cycle", 64, 2, 1.0, False),  # tiny N: many intra-step hazards
    ("er", 300, 7, 1.0, False),    # 4-bit packing, shared-memory counters
    ("er", 300, 40, 0.0, False),   # 8-bit packing, thresholds in global memory
    ("complete", 500, 3, 1.0, False),
    ("complete", 33, 2, 1.0, True),
])
def test_cnvm_native_matches_sequential_oracle(oracle, kind, n, M, alpha, force_global):
    from sponet_b200 import CNVM, CNVMParameters, engine

    rng = np.random.default_rng(3)
    if M == 3:
        r, rt = R3, RT3
    else:
        r, rt = rng.random((M, M)), 0.05 * rng.random((M, M))
    if kind == "complete":
        params = CNVMParameters(num_opinions=M, num_agents=n, r=r, r_tilde=rt, alpha=alpha)
        rp = col = None
    else:
        params = CNVMParameters(num_opinions=M, network=_net(kind, n, 1), r=r, r_tilde=rt, alpha=alpha)
        rp, col = oracle.neighbor_list_to_csr(params.network)
    model = CNVM(params)
    S, R, T = 2, 5, 6
    states = rng.integers(0, M, (S, n)).astype(np.uint8)
    t_eval = np.linspace(0, 4.0, T)
    seed = 0xC0FFEE123456789
    struct, keep = model._struct()
    counters = np.zeros(4, dtype=np.uint64)
    spec = (np.arange(M, dtype=np.int32), 1.0)
    sums = (np.zeros((S, T, M), dtype=np.int64), np.zeros((S, T, M), dtype=np.int64))
    out_x, out_cv = engine.run_native("cnvm", model._graph(), n, struct, states, t_eval, seed, R, 0, R, cv_spec=spec,
                                      want_x=True, want_cv=True, sums=sums, counters=counters,
                                      force_global_state=force_global)
    rate, noise_p = model.event_rates()
    thr_i, thr_n = oracle.prob_to_thr(params.prob_imit), oracle.prob_to_thr(params.prob_noise)
    noise_thr = int(oracle.prob_to_thr(noise_p))
    da = np.asarray(model.degree_alpha, dtype=float)
    if kind != "complete" and not np.all(da == da[0]):
        prob, alias = oracle.build_alias_table(da)
        alias_thr = oracle.prob_to_thr(prob)
    else:
        alias_thr = alias = None
    total_ev = total_acc = 0
    exp_sum = np.zeros((S, T, M), dtype=np.int64)
    exp_sq = np.zeros((S, T, M), dtype=np.int64)
    nbtab, nblg = _nbtab(oracle, model, rp, col)
    for j in range(S):
        for i in range(R):
            chain = j * R + i
            counts = oracle.native_event_counts(seed, chain, 1, t_eval, rate)[0]
            ox, oc, ev, acc = oracle.cnvm_native(states[j], M, rp, col, noise_thr, thr_i, thr_n, alias_thr, alias,
                                                 counts, seed, chain, nbtab=nbtab, nbtab_log2=nblg)
            np.testing.assert_array_equal(out_x[j, i], ox, err_msg=f"states differ for chain ({j},{i})")
            np.testing.assert_array_equal(out_cv[j, i], oc.astype(float))
            exp_sum[j] += oc
            exp_sq[j] += oc * oc
            total_ev += ev
            total_acc += acc
    np.testing.assert_array_equal(sums[0], exp_sum)
    np.testing.assert_array_equal(sums[1], exp_sq)
    assert int(counters[0]) == total_ev and int(counters[1]) == total_acc
    del keep


def test_cnvm_native_sharding_is_invisible(oracle):
    """Runs [0,R) in one call == runs split over two calls (what two GPUs would do)."""
    from sponet_b200 import CNVM, CNVMParameters, engine

    params = CNVMParameters(num_opinions=3, network=_net("er", 400, 2), r=R3, r_tilde=RT3)
    model = CNVM(params)
    states = np.random.default_rng(0).integers(0, 3, (2, 400)).astype(np.uint8)
    t_eval = np.linspace(0, 2.0, 5)
    struct, keep = model._struct()
    spec = (np.array([2, 0], dtype=np.int32), 400.0)
    _, full = engine.run_native("cnvm", model._graph(), 400, struct, states, t_eval, 7, 9, 0, 9, cv_spec=spec, want_cv=True)
    _, a = engine.run_native("cnvm", model._graph(), 400, struct, states, t_eval, 7, 9, 0, 4, cv_spec=spec, want_cv=True)
    _, b = engine.run_native("cnvm", model._graph(), 400, struct, states, t_eval, 7, 9, 4, 9, cv_spec=spec, want_cv=True)
    np.testing.assert_array_equal(full, np.concatenate([a, b], axis=1))
    del keep


def test_cnvm_complete_counts_matches_oracle(oracle):
    from sponet_b200 import CNVM, CNVMParameters, engine

    for M, n in ((2, 1000), (3, 777), (6, 300)):
        rng = np.random.default_rng(M)
        r, rt = (R3, RT3) if M == 3 else (rng.random((M, M)), 0.1 * rng.random((M, M)))
        params = CNVMParameters(num_opinions=M, num_agents=n, r=r, r_tilde=rt)
        model = CNVM(params)
        S, R, T = 2, 40, 7
        c0 = np.stack([np.bincount(rng.integers(0, M, n), minlength=M) for _ in range(S)]).astype(np.int64)
        t_eval = np.linspace(0, 3.0, T)
        struct, keep = model._struct()
        spec = (np.arange(M, dtype=np.int32), 1.0)
        counters = np.zeros(4, dtype=np.uint64)
        _, out = engine.run_native("cnvm_counts", None, n, struct, None, t_eval, 42, R, 0, R, cv_spec=spec, want_cv=True,
                                   counts_init=c0, counters=counters)
        rate, noise_p = model.event_rates()
        thr_i, thr_n = oracle.prob_to_thr(params.prob_imit), oracle.prob_to_thr(params.prob_noise)
        ev_total = 0
        for j in range(S):
            for i in range(R):
                chain = j * R + i
                counts = oracle.native_event_counts(42, chain, 1, t_eval, rate)[0]
                oc, ev = oracle.cnvm_complete_counts_native(c0[j], n, M, int(oracle.prob_to_thr(noise_p)), thr_i, thr_n,
                                                            counts, 42, chain)
                np.testing.assert_array_equal(out[j, i], oc.astype(float))
                ev_total += ev
        assert int(counters[0]) == ev_total
        del keep


@pytest.mark.parametrize("kind,n,force_global", [("er", 500, False), ("ba", 800, False), ("ba", 800, True), ("cycle", 40, False)])
def test_cntm_native_matches_sequential_oracle(oracle, kind, n, force_global):
    import networkx as nx

    from sponet_b200 import CNTM, CNTMParameters, engine

    g = _net(kind, n, 4) if kind != "er" else nx.erdos_renyi_graph(n, 4 / n, seed=4)  # ER keeps isolates
    params = CNTMParameters(g, 1.0, 0.2, 0.4, 0.55)
    model = CNTM(params)
    S, R, T = 2, 4, 6
    states = np.random.default_rng(5).integers(0, 2, (S, n)).astype(np.uint8)
    t_eval = np.linspace(0, 6.0, T)
    seed = 2024
    spec = (np.array([0, 1], dtype=np.int32), 1.0)
    counters = np.zeros(4, dtype=np.uint64)
    out_x, out_cv = engine.run_native("cntm", model._graph(), n, model._struct(), states, t_eval, seed, R, 0, R,
                                      cv_spec=spec, want_x=True, want_cv=True, counters=counters,
                                      force_global_state=force_global)
    rp, col = oracle.neighbor_list_to_csr(model.neighbor_list)
    noise_thr = int(oracle.prob_to_thr(model.noise_prob))
    ev_total = acc_total = 0
    for j in range(S):
        for i in range(R):
            chain = j * R + i
            counts = oracle.native_event_counts(seed, chain, 1, t_eval, model.next_event_rate)[0]
            ox, oc, ev, acc = oracle.cntm_native(states[j], rp, col, noise_thr, 0.4, 0.55, counts, seed, chain)
            np.testing.assert_array_equal(out_x[j, i], ox)
            np.testing.assert_array_equal(out_cv[j, i], oc.astype(float))
            ev_total += ev
            acc_total += acc
    assert int(counters[0]) == ev_total and int(counters[1]) == acc_total


@pytest.mark.parametrize("n,M,alpha", [(40000, 3, 1.0), (40000, 3, 0.5), (70000, 2, 1.0)])
def test_cnvm_pair_mode_matches_single_warp_and_oracle(oracle, n, M, alpha):
    """Two warps per chain (alternating steps, mbarrier hand-over) must give exactly what one warp per
    chain gives -- and one chain is checked against the sequential oracle."""
    import networkx as nx

    from sponet_b200 import CNVM, CNVMParameters, engine

    g = nx.fast_gnp_random_graph(n, 8 / (n - 1), seed=3)
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % n)
    r, rt = (R3, RT3) if M == 3 else (1.0, 0.05)
    params = CNVMParameters(num_opinions=M, network=g, r=r, r_tilde=rt, alpha=alpha)
    model = CNVM(params)
    S, R = 1, 700  # more chains than resident slots: several chains per warp pair
    x0 = np.random.default_rng(1).integers(0, M, (S, n)).astype(np.uint8)
    t_eval = np.array([0.0, 0.004, 0.004001, 0.02, 0.05])  # includes a near-empty interval
    struct, keep = model._struct()
    spec = (np.arange(M, dtype=np.int32), 1.0)
    out = {}
    for wpc in (1, 2, 3, 8):
        counters = np.zeros(4, dtype=np.uint64)
        sums = (np.zeros((S, t_eval.size, M), dtype=np.int64), np.zeros((S, t_eval.size, M), dtype=np.int64))
        _, cv = engine.run_native("cnvm", model._graph(), n, struct, x0, t_eval, 555, R, 0, R, cv_spec=spec,
                                  want_cv=True, sums=sums, counters=counters, warps_per_chain=wpc)
        out[wpc] = (cv, sums, counters)
    for wpc in (2, 3, 8):
        np.testing.assert_array_equal(out[1][0], out[wpc][0])
        np.testing.assert_array_equal(out[1][1][0], out[wpc][1][0])
        np.testing.assert_array_equal(out[1][1][1], out[wpc][1][1])
        np.testing.assert_array_equal(out[1][2][:2], out[wpc][2][:2])  # events, accepted
    # full states of a few chains in pair mode vs the oracle
    xs, cv = engine.run_native("cnvm", model._graph(), n, struct, x0, t_eval, 555, R, 3, 6, cv_spec=spec,
                               want_x=True, want_cv=True, warps_per_chain=2)
    rate, noise_p = model.event_rates()
    rp, col = oracle.neighbor_list_to_csr(params.network)
    da = np.asarray(model.degree_alpha, dtype=float)
    if np.all(da == da[0]):
        alias_thr = alias = None
    else:
        prob, alias = oracle.build_alias_table(da)
        alias_thr = oracle.prob_to_thr(prob)
    nbtab, nblg = _nbtab(oracle, model, rp, col)
    for i in range(3, 6):
        counts = oracle.native_event_counts(555, i, 1, t_eval, rate)[0]
        ox, oc, _, _ = oracle.cnvm_native(x0[0], M, rp, col, int(oracle.prob_to_thr(noise_p)),
                                          oracle.prob_to_thr(params.prob_imit), oracle.prob_to_thr(params.prob_noise),
                                          alias_thr, alias, counts, 555, i, nbtab=nbtab, nbtab_log2=nblg)
        np.testing.assert_array_equal(xs[0, i - 3], ox)
        np.testing.assert_array_equal(cv[0, i - 3], oc.astype(float))
    del keep


def test_headline_size_pair_vs_single_and_conservation(oracle):
    """Full-size workload H (ER N=1e5, nnz 998 782): the two-warps-per-chain schedule (what bench.py runs)
    against one warp per chain -- identical CV trajectories, moment sums and counters -- plus the
    size-independent properties: counts sum to N at every output time, snapshot 0 is the initial
    histogram, events executed = sum of the interval schedule."""
    from sponet_b200 import CNVM, _lib, engine, workloads

    params, x0, rate = workloads.headline_cnvm(100_000)
    model = CNVM(params)
    struct, keep = model._struct()
    n, R, T = 100_000, 1500, 6
    t_eval = np.linspace(0, 0.5, T)
    spec = (np.arange(3, dtype=np.int32), 1.0)
    res = {}
    for wpc in (1, 0):  # 0 = automatic choice: 8 chains x 2 warps per SM
        counters = np.zeros(4, dtype=np.uint64)
        sums = (np.zeros((1, T, 3), dtype=np.int64), np.zeros((1, T, 3), dtype=np.int64))
        _, cv = engine.run_native("cnvm", model._graph(), n, struct, x0[None, :], t_eval, 99, R, 0, R, cv_spec=spec,
                                  want_cv=True, sums=sums, counters=counters, warps_per_chain=wpc)
        res[wpc] = (cv[0], sums, counters)
    np.testing.assert_array_equal(res[1][0], res[0][0])
    np.testing.assert_array_equal(res[1][1][0], res[0][1][0])
    np.testing.assert_array_equal(res[1][1][1], res[0][1][1])
    np.testing.assert_array_equal(res[1][2][:2], res[0][2][:2])
    cv, sums, counters = res[0]
    assert (cv.sum(-1) == n).all()
    np.testing.assert_array_equal(cv[:, 0], np.broadcast_to(np.bincount(x0, minlength=3), (R, 3)))
    np.testing.assert_array_equal(sums[0][0], cv.sum(0).astype(np.int64))
    np.testing.assert_array_equal(sums[1][0], (cv.astype(np.int64) ** 2).sum(0))
    sched = np.empty((R, T - 1), dtype=np.int64)
    _lib.check(_lib.lib().sponet_native_event_counts(99, 0, R, _lib.ptr(t_eval), T, 1.0 / rate, _lib.ptr(sched), None))
    assert int(counters[0]) == int(sched.sum())
    assert abs(sched.sum() / (R * 0.5 * rate) - 1) < 1e-3  # Poisson totals

    # the sequential oracle at the full size of H: agent states AND counts of four chains (the first, two
    # from the middle of the launch -- later waves of their slots -- and the last), bit for bit, in the
    # schedule bench.py runs (automatic: two warps per chain)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    _, noise_p = model.event_rates()
    nbtab, nblg = _nbtab(oracle, model, rp, col)
    assert nblg == 5  # H takes the one-gather neighbour sampler with 32 slots per agent
    for i in (0, 611, 1234, R - 1):
        xs, cvi = engine.run_native("cnvm", model._graph(), n, struct, x0[None, :], t_eval, 99, R, i, i + 1, cv_spec=spec,
                                    want_x=True, want_cv=True)
        np.testing.assert_array_equal(cvi[0, 0], cv[i])  # the same chain alone == inside the full launch
        ox, oc, ev, _ = oracle.cnvm_native(x0, 3, rp, col, int(oracle.prob_to_thr(noise_p)),
                                           oracle.prob_to_thr(params.prob_imit), oracle.prob_to_thr(params.prob_noise),
                                           None, None, sched[i], 99, i, nbtab=nbtab, nbtab_log2=nblg)
        np.testing.assert_array_equal(xs[0, 0], ox, err_msg=f"agent states of chain {i} differ from the sequential oracle")
        np.testing.assert_array_equal(cv[i], oc.astype(float))
        assert ev == int(sched[i].sum())
    del keep


def test_launch_plans_agree_on_h(oracle):
    """The launch policy picks its plan from the chain count: 1 250 chains of H run as ONE wave of the 576-thread
    build (9 chains x 2 warps per SM), 1 500 as a full 16-warp wave plus a separate last-wave launch with more warps
    per chain.  Both must equal the plain one-warp-per-chain launch bit for bit (CVs, moment sums, counters), and a
    chain from the tail of each against the sequential oracle."""
    from sponet_b200 import CNVM, _lib, engine, workloads

    params, x0, rate = workloads.headline_cnvm(100_000)
    model = CNVM(params)
    struct, keep = model._struct()
    n, T = 100_000, 4
    t_eval = np.linspace(0, 0.15, T)
    spec = (np.arange(3, dtype=np.int32), 1.0)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    _, noise_p = model.event_rates()
    nbtab, nblg = _nbtab(oracle, model, rp, col)
    for R in (1250, 1500):
        res = {}
        for wpc in (1, 0):
            counters = np.zeros(4, dtype=np.uint64)
            sums = (np.zeros((1, T, 3), dtype=np.int64), np.zeros((1, T, 3), dtype=np.int64))
            _, cv = engine.run_native("cnvm", model._graph(), n, struct, x0[None, :], t_eval, 7, R, 0, R, cv_spec=spec,
                                      want_cv=True, sums=sums, counters=counters, warps_per_chain=wpc)
            res[wpc] = (cv[0], sums, counters)
        np.testing.assert_array_equal(res[1][0], res[0][0])
        np.testing.assert_array_equal(res[1][1][0], res[0][1][0])
        np.testing.assert_array_equal(res[1][1][1], res[0][1][1])
        np.testing.assert_array_equal(res[1][2][:2], res[0][2][:2])
        sched = np.empty((R, T - 1), dtype=np.int64)
        _lib.check(_lib.lib().sponet_native_event_counts(7, 0, R, _lib.ptr(t_eval), T, 1.0 / rate, _lib.ptr(sched), None))
        for i in (R - 1, R - 40):
            _, oc, _, _ = oracle.cnvm_native(x0, 3, rp, col, int(oracle.prob_to_thr(noise_p)), oracle.prob_to_thr(params.prob_imit),
                                             oracle.prob_to_thr(params.prob_noise), None, None, sched[i], 7, i, nbtab=nbtab,
                                             nbtab_log2=nblg)
            np.testing.assert_array_equal(res[0][0][i], oc.astype(float))
    del keep


def test_config4_sis_ws_1e6_pair_vs_single():
    """BASELINE config 4 shape: SIS-as-CNVM on a Watts-Strogatz ring with N = 1e6 (one chain per SM,
    125 KB of packed state in shared memory).  Pair schedule == single schedule, and the infected share
    follows the SIS drift."""
    from sponet_b200 import CNVM, engine, workloads

    n = 1_000_000
    params, x0, rate = workloads.sis_ws(n, infection_rate=1.0)
    model = CNVM(params)
    struct, keep = model._struct()
    R, T = 200, 3
    t_eval = np.linspace(0, 0.2, T)
    spec = (np.array([1], dtype=np.int32), float(n))
    res = {}
    for wpc in (1, 0, 16):  # 0 = automatic choice (a ring of 8 warps for one chain per SM)
        counters = np.zeros(4, dtype=np.uint64)
        _, cv = engine.run_native("cnvm", model._graph(), n, struct, x0[None, :], t_eval, 11, R, 0, R, cv_spec=spec,
                                  want_cv=True, counters=counters, warps_per_chain=wpc)
        res[wpc] = (cv[0], counters)
    for wpc in (0, 16):
        np.testing.assert_array_equal(res[1][0], res[wpc][0])
        np.testing.assert_array_equal(res[1][1][:2], res[wpc][1][:2])
    cv = res[0][0]
    assert np.allclose(cv[:, 0, 0], 0.3)
    # SIS: infection along edges at rate lambda * (infected neighbours / degree), recovery at rate 1:
    # d rho/dt = lambda rho (1 - rho) - rho = -0.09 at rho = 0.3, lambda = 1 -> the share decays
    assert (cv[:, -1, 0] < 0.3).all() and (cv[:, -1, 0] > 0.27).all()
    assert cv[:, -1, 0].std() < 2e-3  # N = 1e6: runs agree to ~1e-3
    del keep


def test_config3_cntm_ba_1e5_properties():
    """BASELINE config 3 shape: CNTM on Barabasi-Albert N = 1e5, m = 5.  Shared-memory state == global
    state results; counts conserved; snapshot 0 is the initial histogram."""
    from sponet_b200 import CNTM, engine, workloads

    params, x0, rate = workloads.ba_cntm(100_000)
    model = CNTM(params)
    R, T = 600, 4
    t_eval = np.linspace(0, 0.3, T)
    spec = (np.array([0, 1], dtype=np.int32), 1.0)
    res = []
    for force_global in (False, True):
        counters = np.zeros(4, dtype=np.uint64)
        _, cv = engine.run_native("cntm", model._graph(), 100_000, model._struct(), x0[None, :], t_eval, 5, R, 0, R,
                                  cv_spec=spec, want_cv=True, counters=counters, force_global_state=force_global)
        res.append((cv[0], counters))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][1][:2], res[1][1][:2])
    cv = res[0][0]
    assert (cv.sum(-1) == 100_000).all()
    np.testing.assert_array_equal(cv[:, 0], np.broadcast_to(np.bincount(x0, minlength=2), (R, 2)))
    assert abs(int(res[0][1][0]) / (R * 0.3 * rate) - 1) < 2e-3


def test_neighbor_table_construction_and_uniformity(oracle):
    """The table gives every neighbour probability 1/deg (up to 2^-24 per slot); graphs that are not
    eligible (hubs above 64, isolated nodes) keep the CSR path."""
    import networkx as nx

    from sponet_b200 import engine
    from sponet_b200.utils import calculate_neighbor_list

    g = nx.fast_gnp_random_graph(400, 0.03, seed=1)
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % 400)
    nl = calculate_neighbor_list(g)
    gh = engine.GraphHandle(nl)
    tab, lg = library_neighbor_table(gh)
    rp, col = oracle.neighbor_list_to_csr(nl)
    exp, exp_lg = oracle.build_neighbor_table(rp, col)
    assert lg == exp_lg >= 1
    np.testing.assert_array_equal(tab, exp)
    L = 1 << lg
    w0, w1 = (tab & 0xFFFFFFFF).astype(np.int64), (tab >> 32).astype(np.int64)
    thr = ((w0 >> 20) << 12) | (w1 >> 20)
    p0 = np.where((w0 & 0xFFFFF) == (w1 & 0xFFFFF), 1.0, thr / 2.0**24)
    for a in (0, 7, 399):
        prob = {}
        for j in range(L):
            e = a * L + j
            prob[w0[e] & 0xFFFFF] = prob.get(w0[e] & 0xFFFFF, 0.0) + p0[e] / L
            if p0[e] < 1.0:
                prob[w1[e] & 0xFFFFF] = prob.get(w1[e] & 0xFFFFF, 0.0) + (1 - p0[e]) / L
        assert sorted(prob) == sorted(int(v) for v in nl[a])
        assert np.allclose(list(prob.values()), 1.0 / len(nl[a]), atol=2.0**-22)
    star = calculate_neighbor_list(nx.star_graph(80))  # hub of degree 80
    assert library_neighbor_table(engine.GraphHandle(star))[0] is None


@pytest.mark.parametrize("force_global,wpc", [(False, 0), (True, 0), (False, 2)])
def test_edge_cvs_in_kernel_match_oracle_on_the_returned_states(oracle, force_global, wpc):
    """Interfaces / Propensities evaluated by the simulation kernels at every output time
    (a scan of the CV's CSR against the on-chip state) == the reference formulas applied to the states the same
    chains return.  Interface counts and their integer moment sums are exact; the propensity sums are f64 in a
    fixed tree order against the reference's sequential order (rtol 1e-12)."""
    from sponet_b200 import CNVM, CNVMParameters, engine
    from sponet_b200.engine import GraphCache, KernelCV

    n, S, R, T = 1500, 2, 6, 5
    net = _net("er", n, 5)
    r, rt = np.array([[0, 1.3], [0.7, 0]]), np.array([[0, 0.03], [0.05, 0]])
    params = CNVMParameters(num_opinions=2, network=net, r=r, r_tilde=rt, alpha=0.5)
    model = CNVM(params)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    # the CV's own network need not be the model's: Interfaces over a different graph on the same agents
    other = _net("ba", n, 9)
    from sponet_b200.utils import calculate_neighbor_list

    orp, ocol = oracle.neighbor_list_to_csr(calculate_neighbor_list(other))
    cache = GraphCache()
    states = np.random.default_rng(1).integers(0, 2, (S, n)).astype(np.uint8)
    t_eval = np.linspace(0, 3.0, T)
    struct, keep = model._struct()
    kw = dict(force_global_state=force_global, warps_per_chain=wpc)
    xs, _ = engine.run_native("cnvm", model._graph(), n, struct, states, t_eval, 77, R, 0, R, want_x=True, **kw)
    flat = xs.reshape(-1, n)

    spec = KernelCV([], 1.0, kind=KernelCV.INTERFACES, graph=cache.get(calculate_neighbor_list(other)))
    sums = (np.zeros((S, T, 1), dtype=np.int64), np.zeros((S, T, 1), dtype=np.int64))
    _, cv = engine.run_native("cnvm", model._graph(), n, struct, states, t_eval, 77, R, 0, R, cv_spec=spec, want_cv=True,
                              sums=sums, **kw)
    exp = oracle.interfaces(flat, orp, ocol).reshape(S, R, T, 1)
    np.testing.assert_array_equal(cv, exp)
    np.testing.assert_array_equal(sums[0], exp.sum(axis=1).astype(np.int64))
    np.testing.assert_array_equal(sums[1], (exp * exp).sum(axis=1).astype(np.int64))
    spec = KernelCV([], float(other.number_of_edges()), kind=KernelCV.INTERFACES, graph=cache.get(calculate_neighbor_list(other)))
    _, cvn = engine.run_native("cnvm", model._graph(), n, struct, states, t_eval, 77, R, 0, R, cv_spec=spec, want_cv=True, **kw)
    np.testing.assert_array_equal(cvn, exp / other.number_of_edges())

    das = np.diff(rp).astype(float) ** 0.5
    spec = KernelCV([], 1.0, kind=KernelCV.PROPENSITIES, graph=model._graph(), degree_alpha=das,
                    rates=(r[0, 1], r[1, 0], rt[0, 1], rt[1, 0]))
    _, pv = engine.run_native("cnvm", model._graph(), n, struct, states, t_eval, 77, R, 0, R, cv_spec=spec, want_cv=True, **kw)
    np.testing.assert_allclose(pv, oracle.propensities(flat, rp, col, das, r, rt).reshape(S, R, T, 2), rtol=1e-12)
    with pytest.raises(Exception, match="Propensities"):
        engine.run_native("cnvm", model._graph(), n, struct, states, t_eval, 77, R, 0, R, cv_spec=spec,
                          sums=(np.zeros((S, T, 2), dtype=np.int64), np.zeros((S, T, 2), dtype=np.int64)), **kw)
    del keep


def test_edge_cvs_cntm_and_public_api(oracle):
    """Interfaces through the CNTM kernel and through sample_many_runs / sample_moments (host API)."""
    from sponet_b200 import CNTMParameters, CNVMParameters, Interfaces, sample_many_runs, sample_moments
    from sponet_b200.collective_variables import Propensities

    n = 900
    net = _net("ba", n, 4)
    rp, col = oracle.neighbor_list_to_csr([sorted(net.neighbors(i)) for i in range(n)])
    x0 = np.random.default_rng(2).integers(0, 2, n).astype(np.uint8)
    p = CNTMParameters(network=net, r=1.0, r_tilde=0.2, threshold_01=0.5, threshold_10=0.5)
    for normalize in (False, True):
        cvi = Interfaces(net, normalize=normalize)
        _, c = sample_many_runs(p, x0, 2.0, 4, 5, collective_variable=cvi, rng=np.random.default_rng(5))
        _, xs = sample_many_runs(p, x0, 2.0, 4, 5, rng=np.random.default_rng(5))
        exp = oracle.interfaces(xs.reshape(-1, n), rp, col).reshape(5, 4, 1)
        np.testing.assert_array_equal(c, exp / net.number_of_edges() if normalize else exp)
    # CNVM: Propensities(params) in-kernel through the public API, and moments of Interfaces from integer sums
    pv = CNVMParameters(num_opinions=2, network=net, r=1.0, r_tilde=0.05, alpha=1.0)
    cvp = Propensities(pv, normalize=True)
    _, c = sample_many_runs(pv, x0, 1.0, 3, 4, collective_variable=cvp, rng=np.random.default_rng(6))
    _, xs = sample_many_runs(pv, x0, 1.0, 3, 4, rng=np.random.default_rng(6))
    exp = oracle.propensities(xs.reshape(-1, n), rp, col, np.diff(rp).astype(float), pv.r, pv.r_tilde).reshape(4, 3, 2) / n
    np.testing.assert_allclose(c, exp, rtol=1e-12)
    t, mean, var, num = sample_moments(pv, x0, 1.0, 3, 64, rel_tol=1e9, collective_variable=Interfaces(net),
                                       rng=np.random.default_rng(7))
    assert mean.shape == (3, 1) and mean[0, 0] == oracle.interfaces(x0, rp, col)[0, 0] and var[0, 0] == 0.0
    with pytest.raises(ValueError, match="2 opinions"):
        p3 = CNVMParameters(num_opinions=3, network=net, r=R3, r_tilde=RT3)
        sample_many_runs(p3, np.zeros(n, dtype=np.uint8), 1.0, 3, 2, collective_variable=Interfaces(net),
                         rng=np.random.default_rng(0))


def test_crafted_interval_counts_all_schedules(oracle):
    """Interval bookkeeping of the kernels' per-interval loops: injected event counts with empty intervals (leading,
    consecutive, trailing), single events, exactly one / two full steps, one event more or less than a step multiple.
    Every warp schedule (1, 2, 3, 8 warps per chain) and the CNTM kernel against the sequential oracle."""
    from sponet_b200 import CNTM, CNTMParameters, CNVM, CNVMParameters, engine

    n, M, R = 3000, 3, 4
    net = _net("er", n, 11)
    params = CNVMParameters(num_opinions=M, network=net, r=R3, r_tilde=RT3)
    model = CNVM(params)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    base = np.array([0, 0, 1, 31, 32, 33, 0, 0, 64, 63, 65, 1, 0, 97, 128, 5000, 0, 2, 0, 0], dtype=np.int64)
    T = base.size + 1
    counts = np.stack([np.roll(base, 3 * i) for i in range(R)])  # chain i sees the pattern shifted
    counts[1, :] = 0  # a chain without any event
    t_eval = np.linspace(0.0, 1.0, T)
    x0 = np.random.default_rng(8).integers(0, M, (1, n)).astype(np.uint8)
    struct, keep = model._struct()
    spec = (np.arange(M, dtype=np.int32), 1.0)
    rate, noise_p = model.event_rates()
    nbtab, nblg = _nbtab(oracle, model, rp, col)
    exp = [oracle.cnvm_native(x0[0], M, rp, col, int(oracle.prob_to_thr(noise_p)), oracle.prob_to_thr(params.prob_imit),
                              oracle.prob_to_thr(params.prob_noise), None, None, counts[i], 321, i, nbtab=nbtab,
                              nbtab_log2=nblg) for i in range(R)]
    for wpc in (1, 2, 3, 8):
        counters = np.zeros(4, dtype=np.uint64)
        xs, cv = engine.run_native("cnvm", model._graph(), n, struct, x0, t_eval, 321, R, 0, R, cv_spec=spec, want_x=True,
                                   want_cv=True, counters=counters, event_counts=counts, warps_per_chain=wpc)
        for i in range(R):
            np.testing.assert_array_equal(xs[0, i], exp[i][0], err_msg=f"wpc={wpc} chain {i}")
            np.testing.assert_array_equal(cv[0, i], exp[i][1].astype(float))
        assert int(counters[0]) == int(counts.sum()) == sum(e[2] for e in exp)
        assert int(counters[1]) == sum(e[3] for e in exp)
    del keep

    pt = CNTMParameters(network=_net("ba", n, 2), r=1.0, r_tilde=0.3, threshold_01=0.4, threshold_10=0.55)
    mt = CNTM(pt)
    rp, col = oracle.neighbor_list_to_csr(mt.neighbor_list)
    x0 = np.random.default_rng(9).integers(0, 2, (1, n)).astype(np.uint8)
    noise_thr = int(oracle.prob_to_thr(mt.noise_prob))
    counters = np.zeros(4, dtype=np.uint64)
    xs, cv = engine.run_native("cntm", mt._graph(), n, mt._struct(), x0, t_eval, 654, R, 0, R,
                               cv_spec=(np.arange(2, dtype=np.int32), 1.0), want_x=True, want_cv=True, counters=counters,
                               event_counts=counts)
    tot_ev = tot_acc = 0
    for i in range(R):
        ox, oc, ev, acc = oracle.cntm_native(x0[0], rp, col, noise_thr, 0.4, 0.55, counts[i], 654, i)
        np.testing.assert_array_equal(xs[0, i], ox)
        np.testing.assert_array_equal(cv[0, i], oc.astype(float))
        tot_ev += ev
        tot_acc += acc
    assert int(counters[0]) == tot_ev == int(counts.sum()) and int(counters[1]) == tot_acc
