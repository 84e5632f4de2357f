"""CPU: the sequential statement of the device-side Erdos-Renyi generator (oracle_er_graph) and the host generators.

The reference pins no generated graph (its generators are networkx calls on a numpy Generator, sponet/network_generator.py:
26-183), so the bar for the generator is distributional: G(n, p) is defined by independent edges with probability p, which
gives exact moments to test against (tolerances below are 5 sigma of the respective estimator)."""
import numpy as np
import pytest


def _check_simple_graph(row_ptr, col, n):
    deg = np.diff(row_ptr)
    src = np.repeat(np.arange(n), deg)
    assert (src != col).all()  # no self loops
    key = src.astype(np.int64) * n + col
    assert (np.diff(key) > 0).all()  # rows sorted, no duplicate entries
    assert np.array_equal(np.sort(col.astype(np.int64) * n + src), key)  # symmetric
    return deg


@pytest.mark.parametrize("n,p", [(2000, 0.01), (500, 0.3), (50, 1.0), (20000, 5e-4)])
def test_er_oracle_is_a_simple_graph_without_isolates(oracle, n, p):
    rp, col, attempts = oracle.er_graph(n, p, seed=11, c=5, force_no_isolates=True)
    deg = _check_simple_graph(rp, col, n)
    assert deg.min() >= 1 and attempts == 1
    rp2, col2, _ = oracle.er_graph(n, p, seed=11, c=5, force_no_isolates=True)
    assert np.array_equal(rp, rp2) and np.array_equal(col, col2)  # a function of (seed, c)
    rp3, col3, _ = oracle.er_graph(n, p, seed=11, c=6, force_no_isolates=True)
    assert p == 1.0 or not (np.array_equal(rp, rp3) and np.array_equal(col, col3))


def test_er_oracle_edge_statistics(oracle):
    """Edge count ~ Binomial(n(n-1)/2, p); degrees ~ Binomial(n-1, p); the two halves of a row are independent."""
    n, p, G = 1500, 0.004, 40  # mean degree 6: exp(-6) n = 3.7 isolates per graph, patched
    pairs = n * (n - 1) / 2
    edges, degs, iso = [], [], 0
    for c in range(G):
        rp, col, _ = oracle.er_graph(n, p, seed=3, c=c, force_no_isolates=True)
        d = np.diff(rp)
        # the patch adds one edge per (initially) isolated vertex; recover the unpatched edge count bound
        edges.append(rp[-1] / 2)
        degs.append(d)
    edges = np.array(edges)
    exp_iso = n * (1 - p) ** (n - 1)  # each isolate adds at most one edge
    assert abs(edges.mean() - pairs * p - exp_iso / 2) < 5 * np.sqrt(pairs * p * (1 - p) / G) + exp_iso / 2 + 1
    d = np.concatenate(degs).astype(float)
    assert abs(d.mean() - (n - 1) * p) < 5 * np.sqrt((n - 1) * p / d.size) + exp_iso / n
    assert abs(d.var() - (n - 1) * p * (1 - p)) < 0.05 * (n - 1) * p
    # P(degree = k) against the binomial pmf, k = 2 .. 10 (patched isolates only move mass from 0 to 1)
    from math import comb

    for k in range(2, 11):
        pk = comb(n - 1, k) * p**k * (1 - p) ** (n - 1 - k)
        assert abs((d == k).mean() - pk) < 5 * np.sqrt(pk * (1 - pk) / d.size) + 2 * exp_iso / n


def test_er_oracle_rejects_or_repairs_isolates(oracle):
    n, p = 400, 0.012  # mean degree 4.8: a sample has isolates with probability ~0.96
    with pytest.raises(RuntimeError, match="Timeout"):
        oracle.er_graph(n, p, seed=1, c=0, force_no_isolates=False, max_attempts=3)
    rp, col, attempts = oracle.er_graph(n, p, seed=1, c=0, force_no_isolates=True)
    assert np.diff(rp).min() >= 1 and attempts == 1
    _check_simple_graph(rp, col, n)
    # denser: rejection succeeds after a few attempts and returns a graph that needed no repair
    n, p = 300, 0.025  # 300 exp(-7.5) = 0.17 isolates per sample
    got = [oracle.er_graph(n, p, seed=2, c=c, force_no_isolates=False, max_attempts=64) for c in range(30)]
    assert all(np.diff(rp).min() >= 1 for rp, _, _ in got)
    assert 1 < np.mean([a for _, _, a in got]) < 1.6  # 1 / (1 - 0.155) = 1.18 expected attempts
    for (rp, col, a), c in zip(got, range(30)):
        if a == 1:  # an accepted first attempt is the graph the repairing variant returns as well
            rp2, col2, _ = oracle.er_graph(n, p, seed=2, c=c, force_no_isolates=True)
            assert np.array_equal(rp, rp2) and np.array_equal(col, col2)


def test_host_generators_follow_the_reference_protocol():
    nx = pytest.importorskip("networkx")
    from sponet_b200.network_generator import (BarabasiAlbertGenerator, ErdosRenyiGenerator, RandomRegularGenerator,
                                               WattsStrogatzGenerator)

    rng = np.random.default_rng(0)
    er = ErdosRenyiGenerator(200, 0.02, rng=rng, force_no_isolates=True)
    g = er()
    assert g.number_of_nodes() == 200 and nx.number_of_isolates(g) == 0
    assert repr(er) == "Erdos-Renyi random graph with p=0.02 on 200 nodes" and er.abrv() == "ER_p2_N200"
    with pytest.raises(RuntimeError, match="Timeout"):
        ErdosRenyiGenerator(300, 0.005, max_sample_time=0.05, rng=rng)()
    rr = RandomRegularGenerator(50, 4, rng=rng)
    assert all(d == 4 for _, d in rr().degree()) and rr.abrv() == "regular_d4_N50"
    ba = BarabasiAlbertGenerator(60, 3, rng=rng)
    assert ba().number_of_nodes() == 60 and ba.abrv() == "barabasi_m3_N60"
    ws = WattsStrogatzGenerator(40, 4, 0.25, rng=rng)
    assert nx.is_connected(ws()) and ws.abrv() == "watts_k4_p25_N40"
    s1 = ErdosRenyiGenerator(10, 0.5, rng=np.random.default_rng(5)).draw_batch_seed()
    s2 = ErdosRenyiGenerator(10, 0.5, rng=np.random.default_rng(5)).draw_batch_seed()
    assert s1 == s2 and 0 <= s1 < 2**63


def test_ba_oracle_follows_the_networkx_process(oracle):
    """The sequential statement of the device Barabasi-Albert generator: a simple graph with m (n - m) edges whose degree
    distribution is that of networkx.barabasi_albert_graph (same process: star on m + 1 vertices, m distinct
    degree-proportional targets per new vertex) -- compared on pooled samples, 5 sigma of the binomial count per degree."""
    nx = pytest.importorskip("networkx")
    n, m, G = 3000, 4, 12
    ours, theirs = [], []
    for c in range(G):
        rp, col = oracle.ba_graph(n, m, seed=9, c=c)
        deg = _check_simple_graph(rp, col, n)
        assert rp[-1] == 2 * m * (n - m) and deg.min() >= min(m, 1) and deg[m + 1 :].min() >= m
        ours.append(deg)
        theirs.append(np.array([d for _, d in nx.barabasi_albert_graph(n, m, seed=100 + c).degree()]))
    ours, theirs = np.concatenate(ours), np.concatenate(theirs)
    assert ours.sum() == theirs.sum()
    for k in (4, 5, 6, 8, 12, 20):
        pa, pb = (ours == k).mean(), (theirs == k).mean()
        assert abs(pa - pb) < 5 * np.sqrt((pa * (1 - pa) + pb * (1 - pb)) / ours.size) + 1e-4, (k, pa, pb)
    # the hubs are the early vertices in both
    assert ours.reshape(G, n)[:, : m + 1].mean() > 20 * m and theirs.reshape(G, n)[:, : m + 1].mean() > 20 * m
    rp2, col2 = oracle.ba_graph(n, m, seed=9, c=0)
    rp3, col3 = oracle.ba_graph(n, m, seed=9, c=1)
    assert np.array_equal(rp2, oracle.ba_graph(n, m, seed=9, c=0)[0]) and not np.array_equal(col2[: col3.size], col3[: col2.size])
    # m = 1: a tree
    rp, col = oracle.ba_graph(500, 1, seed=1, c=0)
    assert rp[-1] == 2 * 499


def test_ws_oracle_follows_the_networkx_process(oracle):
    """Sequential statement of the device Watts-Strogatz generator against networkx.watts_strogatz_graph: n k / 2 edges, every
    vertex keeps its k / 2 own edges, the share of surviving lattice edges is 1 - p, the degree distribution agrees."""
    nx = pytest.importorskip("networkx")
    n, k, p, G = 2000, 6, 0.2, 10
    ours, theirs, lattice = [], [], []
    for c in range(G):
        rp, col, attempts = oracle.ws_graph(n, k, p, seed=5, c=c)
        deg = _check_simple_graph(rp, col, n)
        assert rp[-1] == n * k and deg.min() >= k // 2 and attempts == 1
        src = np.repeat(np.arange(n), deg)
        lattice.append((np.minimum((src - col) % n, (col - src) % n) <= k // 2).mean())
        ours.append(deg)
        theirs.append(np.array([d for _, d in nx.watts_strogatz_graph(n, k, p, seed=50 + c).degree()]))
    ours, theirs = np.concatenate(ours), np.concatenate(theirs)
    # (a rewired edge lands inside the lattice band by chance with probability ~ k / n)
    assert abs(np.mean(lattice) - (1 - p)) < 5 * np.sqrt(p * (1 - p) / (n * k / 2 * G)) + 2 * k / n
    for d in range(k - 2, k + 4):
        pa, pb = (ours == d).mean(), (theirs == d).mean()
        assert abs(pa - pb) < 5 * np.sqrt((pa * (1 - pa) + pb * (1 - pb)) / ours.size) + 1e-4, (d, pa, pb)
    # p = 0: the lattice itself; a sparse ring that falls apart is drawn again or given up
    rp, col, _ = oracle.ws_graph(50, 4, 0.0, seed=1, c=0)
    assert np.array_equal(col[rp[7] : rp[8]], [5, 6, 8, 9])
    assert any(oracle.ws_graph(400, 2, 0.5, seed=2, c=c)[2] > 1 for c in range(3))
    with pytest.raises(RuntimeError, match="Maximum number of tries"):
        oracle.ws_graph(2000, 2, 0.9, seed=3, c=0, tries=3)


def test_regular_oracle_follows_the_networkx_process(oracle):
    """Sequential statement of the device random-regular generator: simple d-regular graphs; on 6 vertices and d = 3 the 70
    labelled graphs appear with the frequencies networkx.random_regular_graph gives them (the pairing process with restarts is
    not exactly uniform: K_3,3 is over-represented by both) -- chi-square of the two samples, 5 sigma of its distribution."""
    nx = pytest.importorskip("networkx")
    from collections import Counter

    for n, d in ((2000, 4), (501, 6), (64, 63), (50, 1)):
        rp, col, _ = oracle.regular_graph(n, d, seed=3, c=1)
        deg = _check_simple_graph(rp, col, n)
        assert (deg == d).all()
    R = 7000
    ours, theirs = Counter(), Counter()
    rng = np.random.default_rng(1)
    for c in range(R):
        ours[tuple(oracle.regular_graph(6, 3, 99, c)[1])] += 1
        g = nx.random_regular_graph(3, 6, seed=rng)
        theirs[tuple(v for i in range(6) for v in sorted(g[i]))] += 1
    keys = sorted(set(ours) | set(theirs))
    assert len(keys) == 70
    x, y = np.array([ours[k] for k in keys], float), np.array([theirs[k] for k in keys], float)
    chi2 = ((x - y) ** 2 / (x + y)).sum()
    assert chi2 < 69 + 5 * np.sqrt(2 * 69), chi2
    # restarts happen and are counted (dense small graphs fail often)
    assert max(oracle.regular_graph(12, 5, 1, c)[2] for c in range(8)) > 1
