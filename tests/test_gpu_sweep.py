"""GPU: ABI-3 features of the native path -- parameter sweeps (sponet_cnvm_run_sweep), in-kernel CV histograms,
final-state-only output, and the multi-chain steady-state estimator built on them.

Parity bars: a sweep point is BIT-IDENTICAL to a stand-alone call of its parameter set (and that one to the
sequential oracle); histograms equal numpy's histogram of the per-run CV values the same launch returns; the
steady-state estimate lies within the reference's own tolerance of reference-generated fixtures
(tests/golden/make_steady_golden.py).
"""
import numpy as np
import pytest

from conftest import csr_to_neighbor_list, load_golden

pytestmark = pytest.mark.gpu

R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)


def _er(n, seed, d=8):
    import networkx as nx

    g = nx.fast_gnp_random_graph(n, d / (n - 1), seed=seed)
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % n)
    return g


@pytest.mark.parametrize("n,M,alpha", [(3000, 3, 1.0), (3000, 2, 0.5), (40000, 3, 1.0), (900, 6, 1.0), (700, 20, 1.0)])
def test_sweep_point_is_bit_identical_to_a_standalone_call(oracle, n, M, alpha):
    """Every packing width / threshold staging (M <= 4 packed table, M <= 16 shared table, M > 16 global table),
    alias and uniform agent weights, one warp and several warps per chain."""
    from sponet_b200 import CNVM, CNVMParameters, engine
    from sponet_b200.utils import calculate_neighbor_list

    nl = calculate_neighbor_list(_er(n, 5))
    rng = np.random.default_rng(M)
    P, R, T = 4, 37, 5
    plist = []
    for q in range(P):
        r = rng.random((M, M)) * (1 + q)
        rt = 0.05 * rng.random((M, M)) * (1 + q)
        plist.append(CNVMParameters(num_opinions=M, network=nl, r=r, r_tilde=rt, alpha=alpha))
    model = CNVM(plist[0])
    keep = [engine.cnvm_struct(q, model.degree_alpha) for q in plist]
    structs = [k[0] for k in keep]
    from sponet_b200 import _lib

    for st in structs:
        st.degree_alpha = _lib.ptr(keep[0][1][2])
    x0 = rng.integers(0, M, n).astype(np.uint8)
    t_eval = np.linspace(0, 0.8, T)
    spec = (np.arange(M, dtype=np.int32), 1.0)
    seed = 777
    sums = (np.zeros((P, T, M), dtype=np.int64), np.zeros((P, T, M), dtype=np.int64))
    counters = np.zeros(4, dtype=np.uint64)
    cv = engine.run_native_sweep(model._graph(), n, structs, x0[None, :], t_eval, seed, R, 0, R, cv_spec=spec, want_cv=True,
                                 sums=sums, counters=counters)
    assert cv.shape == (P, R, T, M)
    ev = 0
    for q in range(P):
        c2 = np.zeros(4, dtype=np.uint64)
        s2 = (np.zeros((1, T, M), dtype=np.int64), np.zeros((1, T, M), dtype=np.int64))
        _, one = engine.run_native("cnvm", model._graph(), n, structs[q], x0[None, :], t_eval, seed, R, 0, R, cv_spec=spec,
                                   want_cv=True, sums=s2, counters=c2, chain_offset=q * R)
        np.testing.assert_array_equal(cv[q], one[0], err_msg=f"sweep point {q}")
        np.testing.assert_array_equal(sums[0][q], s2[0][0])
        np.testing.assert_array_equal(sums[1][q], s2[1][0])
        ev += int(c2[0])
    assert int(counters[0]) == ev
    # ... and one chain of the last point against the sequential oracle (chain id = q * R + i)
    q, i = P - 1, 11
    pm = CNVM(plist[q])
    rate, noise_p = pm.event_rates()
    rp, col = oracle.neighbor_list_to_csr(nl)
    da = np.asarray(pm.degree_alpha, dtype=float)
    alias_thr = alias = None
    if not np.all(da == da[0]):
        prob, alias = oracle.build_alias_table(da)
        alias_thr = oracle.prob_to_thr(prob)
    nbtab, nblg = oracle.build_neighbor_table(rp, col)
    counts = oracle.native_event_counts(seed, q * R + i, 1, t_eval, rate)[0]
    _, oc, _, _ = oracle.cnvm_native(x0, M, rp, col, int(oracle.prob_to_thr(noise_p)), oracle.prob_to_thr(plist[q].prob_imit),
                                     oracle.prob_to_thr(plist[q].prob_noise), alias_thr, alias, counts, seed, q * R + i,
                                     nbtab=nbtab if nblg >= 0 else None, nbtab_log2=max(nblg, 0))
    np.testing.assert_array_equal(cv[q, i], oc.astype(float))
    del keep


def test_sweep_pieces_do_not_change_the_result():
    """sample_moments_sweep shards the flattened (point, run) range over ranks in up to three pieces per rank; the
    pieces of any world size accumulate to the single-rank sums (what the all-reduce adds up)."""
    import torch

    from sponet_b200 import CNVM, CNVMParameters, OpinionShares, engine, sample_moments_sweep
    from sponet_b200.sample_moments import sweep_pieces
    from sponet_b200.utils import calculate_neighbor_list

    n, P, R, T = 2000, 5, 40, 4
    nl = calculate_neighbor_list(_er(n, 2))
    plist = [CNVMParameters(num_opinions=2, network=nl, r=0.5 + 0.3 * q, r_tilde=0.01 * (q + 1)) for q in range(P)]
    x0 = np.random.default_rng(0).integers(0, 2, n)
    t, mean, var, hist, edges = sample_moments_sweep(plist, x0, 2.0, T, R, OpinionShares(2, normalize=True),
                                                    rng=np.random.default_rng(3), hist_bins=16)
    assert mean.shape == (P, T, 2) and hist.shape == (P, T, 2, 16) and edges.shape == (17,)
    assert (hist.sum(-1) == R).all()
    np.testing.assert_allclose(mean.sum(-1), 1.0, atol=1e-12)
    # the same sums from the pieces of a 3-rank and a 7-rank job
    seed = engine.draw_seed(np.random.default_rng(3))
    model = CNVM(plist[0])
    keep = [engine.cnvm_struct(q, model.degree_alpha) for q in plist]
    from sponet_b200 import _lib

    structs = [k[0] for k in keep]
    for st in structs:
        st.degree_alpha = _lib.ptr(keep[0][1][2])
    spec = OpinionShares(2, normalize=True)._kernel_spec(n)
    d_x0 = torch.as_tensor(x0.astype(np.uint8)[None, :], device="cuda")
    for world in (3, 7):
        sums = torch.zeros((2, P, T, 2), dtype=torch.int64, device="cuda")
        covered = 0
        for rank in range(world):
            for (pb, pe, rb, re) in sweep_pieces(P, R, rank, world):
                engine.run_native_sweep(model._graph(), n, structs[pb:pe], d_x0, t, seed, R, rb, re, cv_spec=spec,
                                        sums=(sums[0, pb:pe], sums[1, pb:pe]), chain_offset=pb * R)
                covered += (pe - pb) * (re - rb)
        assert covered == P * R
        np.testing.assert_allclose(sums[0].cpu().numpy() / R / n, mean, rtol=0, atol=1e-15)
    del keep


@pytest.mark.parametrize("kind", ["cnvm", "cnvm_complete", "cnvm_counts", "cntm"])
def test_histogram_and_final_state_only(kind):
    import networkx as nx

    from sponet_b200 import CNTM, CNTMParameters, CNVM, CNVMParameters, engine
    from sponet_b200.utils import calculate_neighbor_list

    n, R, T, bins = 1500, 300, 6, 25
    rng = np.random.default_rng(4)
    g = _er(n, 9)
    if kind == "cntm":
        model = CNTM(CNTMParameters(network=g, r=1.0, r_tilde=0.2, threshold_01=0.4, threshold_10=0.5))
        struct, keep, M, graph = model._struct(), None, 2, model._graph()
    else:
        net = calculate_neighbor_list(g) if kind == "cnvm" else None
        model = CNVM(CNVMParameters(num_opinions=3, network=net, num_agents=None if net is not None else n, r=R3, r_tilde=RT3))
        (struct, keep), M = model._struct(), 3
        graph = model._graph()
    x0 = rng.integers(0, M, (2, n)).astype(np.uint8)
    t_eval = np.linspace(0, 1.5, T)
    spec = (np.arange(M, dtype=np.int32), 1.0)
    lo, width = engine.Histogram.for_counts(bins, n)
    h = engine.Histogram(bins, lo, width, np.zeros((2, T, M, bins), dtype=np.uint64))
    if kind == "cnvm_counts":
        c0 = np.stack([np.bincount(s, minlength=M) for s in x0]).astype(np.int64)
        _, cv = engine.run_native("cnvm_counts", None, n, struct, None, t_eval, 5, R, 0, R, cv_spec=spec, want_cv=True,
                                  counts_init=c0, hist=h)
    else:
        k = "cntm" if kind == "cntm" else "cnvm"
        full, cv = engine.run_native(k, graph, n, struct, x0, t_eval, 5, R, 0, R, cv_spec=spec, want_x=True, want_cv=True, hist=h)
        last, cv2 = engine.run_native(k, graph, n, struct, x0, t_eval, 5, R, 0, R, cv_spec=spec, want_x=True, want_cv=True,
                                      final_state_only=True)
        assert last.shape == (2, R, n)
        np.testing.assert_array_equal(last, full[:, :, -1])
        np.testing.assert_array_equal(cv2, cv)
    exp = np.zeros_like(h.out)
    b = np.minimum((cv.astype(np.int64) - lo) // width, bins - 1)  # [S, R, T, M]
    for s in range(2):
        for k_ in range(T):
            for m in range(M):
                exp[s, k_, m] = np.bincount(b[s, :, k_, m], minlength=bins)
    np.testing.assert_array_equal(h.out, exp)
    assert (h.out.sum(-1) == R).all()
    del keep


def test_sample_histogram_api():
    from sponet_b200 import CNVMParameters, OpinionShares, sample_histogram, sample_many_runs
    from sponet_b200.utils import calculate_neighbor_list

    n, R, T = 1000, 500, 4
    nl = calculate_neighbor_list(_er(n, 1))
    p = CNVMParameters(num_opinions=3, network=nl, r=R3, r_tilde=RT3)
    x0 = np.random.default_rng(0).integers(0, 3, n)
    cv = OpinionShares(3, normalize=True)
    t, hist, edges, mean, var = sample_histogram(p, x0, 2.0, T, R, cv, bins=20, rng=np.random.default_rng(8))
    _, c = sample_many_runs(p, x0, 2.0, T, R, collective_variable=cv, rng=np.random.default_rng(8))
    assert hist.shape == (T, 3, 20) and (hist.sum(-1) == R).all()
    np.testing.assert_allclose(mean, c.mean(0), atol=1e-12)
    np.testing.assert_allclose(var, c.var(0), atol=1e-12)
    for k in range(T):
        for m in range(3):
            np.testing.assert_array_equal(hist[k, m], np.histogram(c[:, k, m], bins=edges)[0])


@pytest.mark.parametrize("name", ["steady_cnvm_er400", "steady_cntm_er400"])
def test_multi_chain_steady_state_matches_reference(name):
    """estimate_steady_state with 128 chains side by side against the mean of 8 runs of the unmodified reference
    estimator (same tol_abs = 0.005): within 2 * tol_abs (the reference's own runs scatter by 0.005-0.007)."""
    import networkx as nx

    from sponet_b200 import CNTMParameters, CNVMParameters, OpinionShares, estimate_steady_state

    g = load_golden(name)
    nl = csr_to_neighbor_list(g["row_ptr"], g["col"])
    M = int(g["num_opinions"])
    if str(g["model"]) == "cntm":
        net = nx.Graph()
        net.add_nodes_from(range(len(nl)))
        net.add_edges_from((i, int(j)) for i, nb in enumerate(nl) for j in nb if i < j)
        params = CNTMParameters(net, float(g["r"]), float(g["r_tilde"]), float(g["threshold_01"]), float(g["threshold_10"]))
    else:
        params = CNVMParameters(num_opinions=M, network=nl, r=g["r"], r_tilde=g["r_tilde"])
    tol = float(g["tol_abs"])
    est = estimate_steady_state(params, OpinionShares(M, normalize=True), tol_abs=tol, rng=np.random.default_rng(1), verbose=False)
    assert est.shape == (M,)
    assert abs(est.sum() - 1.0) < 1e-12
    assert np.abs(est - g["mean"]).max() <= 2 * tol, (est, g["mean"])
    # a second seed agrees with the first within the tolerance as well (and is not identical)
    est2 = estimate_steady_state(params, OpinionShares(M, normalize=True), tol_abs=tol, rng=np.random.default_rng(2), verbose=False)
    assert np.abs(est - est2).max() <= 2 * tol and not np.array_equal(est, est2)
