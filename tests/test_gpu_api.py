"""GPU: API-level behaviour mirroring the reference's own tests (shapes, dtypes, determinism,
absorbing states, CV known answers) -- tests/tests_cnvm/test_cnvm_simulate.py,
tests/tests_cntm/test_cntm_simulate.py, tests/test_multiprocessing.py, tests/test_sample_moments.py,
tests/test_steady_state.py, tests/tests_collective_variables/*.py of the reference."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)


def _networks():
    import networkx as nx

    return [None, nx.barabasi_albert_graph(100, 2, seed=1), nx.cycle_graph(100)]


@pytest.mark.parametrize("mode", ["native", "replay"])
@pytest.mark.parametrize("net_id", [0, 1, 2])
def test_cnvm_simulate_contract(mode, net_id):
    from sponet_b200 import CNVM, CNVMParameters

    net = _networks()[net_id]
    params = CNVMParameters(num_opinions=3, num_agents=100, network=net, r=R3, r_tilde=RT3)
    model = CNVM(params)
    x0 = np.random.default_rng(0).integers(0, 3, 100)
    t, x = model.simulate(10.0, x0, 11, rng=np.random.default_rng(1), mode=mode)
    assert t.shape == (11,) and x.shape == (11, 100) and x.dtype == np.uint8
    assert t[0] == 0 and 9.9 < t[-1] < 10.1 and np.allclose(t, np.linspace(0, 10, 11), atol=0.01)
    assert (x[0] == x0).all() and set(np.unique(x)) <= {0, 1, 2}
    t2, x2 = model.simulate(10.0, x0, 11, rng=np.random.default_rng(1), mode=mode)  # same seed -> same result
    assert (t == t2).all() and (x == x2).all()
    t3, x3 = model.simulate(10.0, x0, [1.0, 2.5, 7.0], rng=np.random.default_rng(2), mode=mode)
    assert t3.shape == (3,) and x3.shape == (3, 100) and np.allclose(t3, [1.0, 2.5, 7.0], atol=0.01)
    tr, xr = model.simulate(5.0, None, 3, rng=np.random.default_rng(3), mode=mode)  # random initial state
    assert xr.shape == (3, 100)


def test_store_all_is_concise_and_fine_grid_has_duplicates():
    from sponet_b200 import CNVM, CNVMParameters
    from sponet_b200.utils import mask_subsequent_duplicates

    params = CNVMParameters(num_opinions=3, num_agents=100, r=R3, r_tilde=RT3)
    model = CNVM(params)
    x0 = np.random.default_rng(0).integers(0, 3, 100)
    t, x = model.simulate(1.0, x0, rng=np.random.default_rng(1))
    assert t[0] == 0 and x.shape == (t.shape[0], 100) and (np.diff(t) > 0).all()
    assert mask_subsequent_duplicates(x).all()  # every stored snapshot differs from its predecessor
    t, x = model.simulate(0.01, x0, 2001, rng=np.random.default_rng(1))  # grid finer than the event spacing
    assert not mask_subsequent_duplicates(x).all()


@pytest.mark.parametrize("M,dtype", [(2, np.uint8), (10, np.uint8), (257, np.uint16)])
def test_output_dtype(M, dtype):
    from sponet_b200 import CNVM, CNVMParameters, sample_many_runs

    params = CNVMParameters(num_opinions=M, num_agents=50, r=1, r_tilde=0.1)
    t, x = CNVM(params).simulate(1.0, None, 3, rng=np.random.default_rng(0))
    assert x.dtype == dtype and x.max() < M
    x0 = np.random.default_rng(1).integers(0, M, 50)
    _, xs = sample_many_runs(params, x0, 1.0, 3, 2, rng=np.random.default_rng(2))
    assert xs.dtype == dtype and xs.shape == (2, 3, 50)


@pytest.mark.parametrize("mode", ["native", "replay"])
def test_absorbing_state(mode):
    """r = [[0,1],[0,0]], r_tilde = 0: opinion 0 dies out and the all-1 state is kept."""
    from sponet_b200 import CNVM, CNVMParameters

    params = CNVMParameters(num_opinions=2, num_agents=50, r=np.array([[0, 1], [0, 0]]), r_tilde=0)
    x0 = np.array([0] * 25 + [1] * 25)
    t, x = CNVM(params).simulate(200.0, x0, 21, rng=np.random.default_rng(5), mode=mode)
    assert (x[-1] == 1).all()
    first = int(np.argmax((x == 1).all(axis=1)))
    assert (x[first:] == 1).all()


@pytest.mark.parametrize("mode", ["native", "replay"])
def test_cntm_simulate_contract(mode):
    import networkx as nx

    from sponet_b200 import CNTM, CNTMParameters

    g = nx.erdos_renyi_graph(100, 0.1, seed=123)
    model = CNTM(CNTMParameters(g, 1, 0.2, 0.5, 0.5))
    x0 = np.random.default_rng(0).integers(0, 2, 100)
    t, x = model.simulate(10.0, x0, 11, rng=np.random.default_rng(1), mode=mode)
    assert t.shape == (11,) and x.shape == (11, 100) and x.dtype == np.uint8 and (x[0] == x0).all()
    assert np.allclose(t, np.linspace(0, 10, 11), atol=0.05)
    t2, x2 = model.simulate(10.0, x0, 11, rng=np.random.default_rng(1), mode=mode)
    assert (x == x2).all() and (t == t2).all()
    # thresholds 0 / 1 and no noise: every evaluated node with >= 0 share of ones flips to 1 and stays
    model = CNTM(CNTMParameters(nx.complete_graph(30), 1, 0.0, 0.0, 1.1))
    t, x = model.simulate(50.0, np.zeros(30, dtype=int), 6, rng=np.random.default_rng(2), mode=mode)
    assert (x[-1] == 1).all()
    ta, xa = model.simulate(2.0, np.zeros(30, dtype=int), rng=np.random.default_rng(3))  # store-all
    assert xa.shape[1] == 30 and (np.diff(xa.sum(1)) == 1).all()


def test_sample_many_runs_shapes_and_determinism():
    import networkx as nx

    from sponet_b200 import CNTMParameters, CNVMParameters, OpinionShares, sample_many_runs

    params = CNVMParameters(num_opinions=3, network=nx.barabasi_albert_graph(100, 2, seed=1), r=R3, r_tilde=RT3)
    states = np.random.default_rng(0).integers(0, 3, (4, 100))
    t, x = sample_many_runs(params, states, 2.0, 5, 6, rng=np.random.default_rng(1))
    assert t.shape == (5,) and x.shape == (4, 6, 5, 100) and x.dtype == np.uint8
    assert (x[:, :, 0] == states[:, None, :]).all()
    t, x1 = sample_many_runs(params, states[0], 2.0, 5, 6, n_jobs=2, rng=np.random.default_rng(1))
    assert x1.shape == (6, 5, 100)  # 1-D initial state drops the leading axis
    cv = OpinionShares(3, normalize=True, idx_to_return=[0, 2])
    t, c = sample_many_runs(params, states, 2.0, [0.5, 1.0, 2.0], 6, collective_variable=cv, rng=np.random.default_rng(1))
    assert c.shape == (4, 6, 3, 2) and c.dtype == np.float64 and (c >= 0).all() and (c <= 1).all()
    _, c2 = sample_many_runs(params, states, 2.0, [0.5, 1.0, 2.0], 6, n_jobs=-1, collective_variable=cv,
                             rng=np.random.default_rng(1))
    assert (c == c2).all()  # native mode: same seed -> same output, whatever n_jobs says
    # in-kernel CV == CV evaluated on the stored states of the same chains
    _, xs = sample_many_runs(params, states, 2.0, [0.5, 1.0, 2.0], 6, rng=np.random.default_rng(1))
    np.testing.assert_array_equal(cv(xs.reshape(-1, 100)).reshape(c.shape), c)
    with pytest.raises(ValueError, match="Parameters not valid"):
        sample_many_runs(object(), states, 1.0, 2, 1)
    p2 = CNTMParameters(nx.erdos_renyi_graph(100, 0.1, seed=3), 1, 0.2, 0.5, 0.5)
    _, xt = sample_many_runs(p2, states[:2] % 2, 2.0, 4, 3, collective_variable=OpinionShares(2), rng=np.random.default_rng(2))
    assert xt.shape == (2, 3, 4, 2) and (xt.sum(-1) == 100).all()


def test_complete_graph_count_only_path_agrees_with_agent_level_statistics():
    from sponet_b200 import CNVM, CNVMParameters, OpinionShares, engine, sample_many_runs

    params = CNVMParameters(num_opinions=3, num_agents=300, r=R3, r_tilde=RT3)
    x0 = np.random.default_rng(0).integers(0, 3, 300)
    _, c = sample_many_runs(params, x0, 3.0, 4, 20000, collective_variable=OpinionShares(3), rng=np.random.default_rng(1))
    model = CNVM(params)
    struct, keep = model._struct()
    _, a = engine.run_native("cnvm", None, 300, struct, x0[None].astype(np.uint8), np.linspace(0, 3, 4), 99, 20000, 0,
                             20000, cv_spec=(np.arange(3, dtype=np.int32), 1.0), want_cv=True)
    a = a[0]
    z = np.abs(c.mean(0) - a.mean(0))[1:] / np.sqrt(c.var(0)[1:] / 20000 + a.var(0)[1:] / 20000)
    assert z.max() < 4.75


def test_generic_cv_paths():
    import networkx as nx

    from sponet_b200 import (CNVMParameters, CompositeCollectiveVariable, DegreeWeightedOpinionShares, OpinionShares,
                             OpinionSharesByDegree, sample_many_runs)

    g = nx.barabasi_albert_graph(100, 2, seed=1)
    params = CNVMParameters(num_opinions=3, network=g, r=R3, r_tilde=RT3)
    x0 = np.random.default_rng(0).integers(0, 3, 100)
    w = np.random.default_rng(1).random(100)
    cvs = [OpinionShares(3, weights=w), DegreeWeightedOpinionShares(3, g, normalize=True), OpinionSharesByDegree(3, g),
           CompositeCollectiveVariable([OpinionShares(3), OpinionShares(3, True, w, 1)])]
    _, xs = sample_many_runs(params, x0, 1.0, 3, 4, rng=np.random.default_rng(5))
    for cv in cvs:
        _, c = sample_many_runs(params, x0, 1.0, 3, 4, collective_variable=cv, rng=np.random.default_rng(5))
        assert c.shape == (4, 3, cv.dimension)
        np.testing.assert_allclose(c, cv(xs.reshape(-1, 100)).reshape(c.shape), rtol=1e-13, atol=0)

    class HostOnlyCV:  # a user-defined CV that only understands numpy
        dimension = 1

        def __call__(self, x):
            return np.asarray(x, dtype=float).sum(axis=-1, keepdims=True)

    _, c = sample_many_runs(params, x0, 1.0, 3, 4, collective_variable=HostOnlyCV(), rng=np.random.default_rng(5))
    np.testing.assert_array_equal(c[..., 0], xs.sum(-1))


def test_collective_variables_known_answers():
    """Values produced by the reference's own classes (fixture cv_known_answers)."""
    import networkx as nx

    from sponet_b200 import DegreeWeightedOpinionShares, OpinionShares, OpinionSharesByDegree

    g = load_golden("cv_known_answers")
    x, w = g["x"], g["weights"]
    net = nx.Graph()
    net.add_nodes_from(range(100))
    net.add_edges_from((i, int(j)) for i in range(100) for j in g["col"][g["row_ptr"][i]:g["row_ptr"][i + 1]] if i < j)
    assert [d for _, d in net.degree()] == g["degrees"].tolist()
    np.testing.assert_array_equal(OpinionShares(3)(x), g["shares"])
    np.testing.assert_array_equal(OpinionShares(3, normalize=True)(x), g["shares_norm"])
    # float weights: fixed-tree summation on the GPU vs sequential bincount -> 1e-13 relative
    np.testing.assert_allclose(OpinionShares(3, weights=w)(x), g["shares_w"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(OpinionShares(3, True, w, [2, 0])(x), g["shares_w_norm"], rtol=1e-13, atol=1e-13)
    np.testing.assert_array_equal(DegreeWeightedOpinionShares(3, net)(x), g["shares_deg"])  # integer weights: exact
    np.testing.assert_array_equal(DegreeWeightedOpinionShares(3, net, True, 1)(x), g["shares_deg_norm"])
    np.testing.assert_array_equal(OpinionSharesByDegree(3, net)(x), g["by_degree"])
    np.testing.assert_array_equal(OpinionSharesByDegree(3, net, True, [0, 2])(x), g["by_degree_norm"])
    np.testing.assert_array_equal(OpinionShares(3)(x[0]), g["shares"][0])  # 1-D input -> 1-D output
    # the reference's hand-written cases (tests_collective_variables/test_opinion_shares.py)
    xs = np.array([[0, 1, 0, 1, 2, 2, 0, 0, 1, 0], [1, 1, 1, 1, 1, 1, 1, 1, 1, 1]])
    np.testing.assert_array_equal(OpinionShares(3)(xs), [[5, 3, 2], [0, 10, 0]])
    np.testing.assert_array_equal(OpinionShares(3, idx_to_return=0)(xs), [[5], [0]])
    star = nx.star_graph(9)  # centre has degree 9
    np.testing.assert_array_equal(DegreeWeightedOpinionShares(3, star)(xs[0]), [9 + 4, 3, 2])


def test_sample_moments_and_steady_state():
    import time

    import networkx as nx

    from sponet_b200 import CNVMParameters, OpinionShares, estimate_steady_state, sample_many_runs, sample_moments

    params = CNVMParameters(num_opinions=3, network=nx.barabasi_albert_graph(100, 2, seed=1), r=R3, r_tilde=RT3)
    x0 = np.random.default_rng(0).integers(0, 3, 100)
    cv = OpinionShares(3, normalize=True)
    t, mean, var, n = sample_moments(params, x0, 2.0, 5, 500, rel_tol=0.05, collective_variable=cv, rng=np.random.default_rng(1))
    assert t.shape == (5,) and mean.shape == (5, 3) and var.shape == (5, 3) and n % 500 == 0
    np.testing.assert_allclose(mean[0], np.bincount(x0, minlength=3) / 100)
    np.testing.assert_allclose(mean.sum(1), 1.0)
    # fused integer moments == moments of the per-run values of the same chains (same seed)
    _, c = sample_many_runs(params, x0, 2.0, 5, 500, collective_variable=cv, rng=np.random.default_rng(7))
    _, m1, v1, n1 = sample_moments(params, x0, 2.0, 5, 500, rel_tol=1e9, collective_variable=cv, rng=np.random.default_rng(7))
    assert n1 == 500
    np.testing.assert_allclose(m1, c.mean(0), rtol=1e-12)
    np.testing.assert_allclose(v1, c.var(0), rtol=1e-9, atol=1e-15)
    # no CV: moments of the agent states; timeout honoured
    start = time.time()
    t, mean, var, n = sample_moments(params, x0, 1.0, 3, 50, rel_tol=1e-9, timeout_seconds=2, rng=np.random.default_rng(2))
    assert time.time() - start < 20 and mean.shape == (3, 100)
    p2 = CNVMParameters(num_opinions=2, num_agents=100, r=1, r_tilde=0.2)
    est = estimate_steady_state(p2, OpinionShares(2, normalize=True), tol_abs=0.05, rng=np.random.default_rng(3), verbose=False)
    assert est.shape == (2,) and abs(est.sum() - 1) < 1e-12 and abs(est[0] - 0.5) < 0.2


def test_interfaces_and_propensities_known_answers():
    """values from the reference's own Interfaces / Propensities classes."""
    import networkx as nx

    from sponet_b200 import CNVMParameters, Interfaces
    from sponet_b200.collective_variables import Propensities

    g = load_golden("cv2_known_answers")
    rp, col, x = g["row_ptr"], g["col"], g["x"]
    nl = [np.ascontiguousarray(col[rp[i]:rp[i + 1]]) for i in range(120)]
    net = nx.Graph()
    net.add_nodes_from(range(120))
    net.add_edges_from((i, int(j)) for i in range(120) for j in nl[i] if i < j)
    assert net.number_of_edges() == int(g["num_edges"])
    np.testing.assert_array_equal(Interfaces(net)(x), g["interfaces"])  # integer counts: exact
    np.testing.assert_array_equal(Interfaces(net, True)(x), g["interfaces_norm"])
    np.testing.assert_array_equal(Interfaces(net)(x[0]), g["interfaces_1d"])
    with pytest.raises(ValueError, match="2 opinions"):
        Interfaces(net)(x + 1)
    for alpha in (1.0, 0.5, 0.0):
        p = CNVMParameters(num_opinions=2, network=nl, r=g["r"], r_tilde=g["r_tilde"], alpha=alpha)
        # float sums in a fixed tree order vs the reference's sequential order: 1e-13 relative
        np.testing.assert_allclose(Propensities(p)(x), g[f"prop_net_a{alpha}"], rtol=1e-13)
        np.testing.assert_allclose(Propensities(p, True)(x), g[f"prop_net_norm_a{alpha}"], rtol=1e-13)
        pc = CNVMParameters(num_opinions=2, num_agents=120, r=g["r"], r_tilde=g["r_tilde"], alpha=alpha)
        np.testing.assert_allclose(Propensities(pc)(x), g[f"prop_complete_a{alpha}"], rtol=1e-13)
    # usable as collective variable of the ensemble driver (evaluated from device-resident slabs)
    from sponet_b200 import sample_many_runs

    p = CNVMParameters(num_opinions=2, network=nl, r=1.0, r_tilde=0.02)
    _, c = sample_many_runs(p, x[0], 1.0, 3, 4, collective_variable=Interfaces(net), rng=np.random.default_rng(0))
    _, xs = sample_many_runs(p, x[0], 1.0, 3, 4, rng=np.random.default_rng(0))
    np.testing.assert_array_equal(c, Interfaces(net)(xs.reshape(-1, 120)).reshape(4, 3, 1))


def test_full_trajectories_are_chunk_invariant(monkeypatch):
    """cv=None output is produced in run chunks (bounded device staging): same result for any chunk size."""
    import networkx as nx

    from sponet_b200 import CNVMParameters, sample_many_runs
    from sponet_b200 import multiprocessing as smp

    params = CNVMParameters(num_opinions=3, network=nx.barabasi_albert_graph(100, 2, seed=1), r=R3, r_tilde=RT3)
    states = np.random.default_rng(0).integers(0, 3, (2, 100))
    _, a = sample_many_runs(params, states, 1.0, 4, 7, rng=np.random.default_rng(3))
    monkeypatch.setattr(smp, "_SLAB_BYTES", 2 * 4 * 100 * 3)  # three runs per chunk
    _, b = sample_many_runs(params, states, 1.0, 4, 7, rng=np.random.default_rng(3))
    np.testing.assert_array_equal(a, b)


class _RingGenerator:
    """Minimal stand-in for a sponet NetworkGenerator: `num_agents` + `__call__() -> nx.Graph`
    (sponet/network_generator.py protocol).  Each call returns a differently rewired ring."""

    def __init__(self, n, seed):
        self.num_agents = n
        self.rng = np.random.default_rng(seed)
        self.calls = 0

    def __call__(self):
        import networkx as nx

        self.calls += 1
        return nx.connected_watts_strogatz_graph(self.num_agents, 4, 0.3, seed=int(self.rng.integers(1 << 30)))


def test_network_generator_resamples_per_run():
    """CNVMParameters(network_generator=...): a new network per simulate() call and per run of
    sample_many_runs (sponet/cnvm/model.py:95-96, sponet/multiprocessing.py:36)."""
    from sponet_b200 import CNVM, CNVMParameters, OpinionShares, sample_many_runs

    gen = _RingGenerator(60, seed=1)
    params = CNVMParameters(num_opinions=2, network_generator=gen, r=1.0, r_tilde=0.05)
    assert params.num_agents == 60 and params.network is None
    model = CNVM(params)
    x0 = np.random.default_rng(0).integers(0, 2, 60)
    t, x = model.simulate(2.0, x0, 5, rng=np.random.default_rng(1))
    assert gen.calls == 1 and x.shape == (5, 60) and len(params.network) == 60
    first = [nb.copy() for nb in params.network]
    model.simulate(2.0, x0, 5, rng=np.random.default_rng(1))
    assert gen.calls == 2 and any(not np.array_equal(a, b) for a, b in zip(first, params.network))
    with pytest.raises(ValueError, match="NetworkGenerator"):
        params.set_network(None)
    _, c = sample_many_runs(params, x0, 1.0, 3, 4, collective_variable=OpinionShares(2), rng=np.random.default_rng(2))
    assert c.shape == (4, 3, 2) and gen.calls == 6 and (c.sum(-1) == 60).all()


def test_sample_moments_resamples_the_network_per_run():
    """sample_moments with a NetworkGenerator must draw a fresh network for every run like the reference
    (sample_moments.py:96 -> sample_many_runs -> simulate -> update_network), not run the fused path on one
    (complete / stale) network: its sums equal those of sample_many_runs on an identical generator."""
    from sponet_b200 import CNVMParameters, OpinionShares, sample_many_runs, sample_moments

    n, R, T = 60, 12, 4
    x0 = np.random.default_rng(0).integers(0, 2, n)
    gen = _RingGenerator(n, seed=3)
    params = CNVMParameters(num_opinions=2, network_generator=gen, r=1.0, r_tilde=0.05)
    t, mean, var, ns = sample_moments(params, x0, 2.0, T, R, rel_tol=1e9, collective_variable=OpinionShares(2),
                                      rng=np.random.default_rng(4))
    assert gen.calls == R and ns == R
    ref = CNVMParameters(num_opinions=2, network_generator=_RingGenerator(n, seed=3), r=1.0, r_tilde=0.05)
    _, c = sample_many_runs(ref, x0, 2.0, T, R, collective_variable=OpinionShares(2), rng=np.random.default_rng(4))
    np.testing.assert_allclose(mean, c.mean(0), rtol=0, atol=1e-12)
    np.testing.assert_allclose(var, (c**2).mean(0) - c.mean(0) ** 2, rtol=0, atol=1e-9)


def test_out_of_range_initial_opinions_are_rejected():
    """An opinion >= num_opinions would be packed into the neighbouring agent's bits: ValueError from
    the Python layer, SPONET_ERR_INVALID from the C ABI for host buffers."""
    from sponet_b200 import CNVM, CNVMParameters, OpinionShares, engine, sample_many_runs, workloads

    params, x0, _ = workloads.headline_cnvm(300)
    bad = x0.copy()
    bad[7] = 3
    with pytest.raises(ValueError, match="opinions in"):
        CNVM(params).simulate(1.0, bad, 3, rng=np.random.default_rng(0))
    with pytest.raises(ValueError, match="opinions in"):
        sample_many_runs(params, bad, 1.0, 3, 2, collective_variable=OpinionShares(3), rng=np.random.default_rng(0))
    model = CNVM(params)
    struct, keep = model._struct()
    with pytest.raises(ValueError, match="not an opinion"):
        engine.run_native("cnvm", model._graph(), 300, struct, bad[None, :], np.linspace(0, 1, 3), 1, 2, 0, 2,
                          cv_spec=(np.arange(3, dtype=np.int32), 1.0), want_cv=True)
    del keep


def test_model_cache_notices_in_place_rewiring():
    """sample_many_runs keeps the model (CSR on the GPU) per parameter object; an adjacency mutated in place
    keeps its id and its edge count, so the cache keys on a content hash as well."""
    import networkx as nx

    from sponet_b200 import CNTMParameters, OpinionShares, sample_many_runs

    g = nx.cycle_graph(40)
    params = CNTMParameters(network=g, r=1.0, r_tilde=0.1, threshold_01=0.5, threshold_10=0.5)
    x0 = np.random.default_rng(0).integers(0, 2, 40)
    kw = dict(collective_variable=OpinionShares(2), rng=None)
    run = lambda: sample_many_runs(params, x0, 3.0, 4, 50, collective_variable=OpinionShares(2),  # noqa: E731
                                   rng=np.random.default_rng(5))[1]
    a = run()
    np.testing.assert_array_equal(a, run())
    nx.double_edge_swap(g, nswap=15, max_tries=1000, seed=1)  # same node / edge counts, same object
    b = run()
    fresh = CNTMParameters(network=g.copy(), r=1.0, r_tilde=0.1, threshold_01=0.5, threshold_10=0.5)
    c = sample_many_runs(fresh, x0, 3.0, 4, 50, collective_variable=OpinionShares(2), rng=np.random.default_rng(5))[1]
    np.testing.assert_array_equal(b, c)
    assert not np.array_equal(a, b)
    del kw


def test_network_generator_native_chains_match_oracle(oracle):
    """Native mode with a NetworkGenerator: pair c = (state j, run i) is one chain on the c-th network of the
    generator, keyed by chain id c -- checked against the sequential oracle on the same networks."""
    from sponet_b200 import CNVM, CNVMParameters, OpinionShares, engine, sample_many_runs
    from sponet_b200.utils import calculate_neighbor_list

    n, S, R, T = 200, 2, 3, 4
    params = CNVMParameters(num_opinions=2, network_generator=_RingGenerator(n, seed=5), r=1.0, r_tilde=0.05)
    states = np.random.default_rng(0).integers(0, 2, (S, n)).astype(np.uint8)
    _, c = sample_many_runs(params, states, 3.0, T, R, collective_variable=OpinionShares(2), rng=np.random.default_rng(9))
    _, x = sample_many_runs(CNVMParameters(num_opinions=2, network_generator=_RingGenerator(n, seed=5), r=1.0, r_tilde=0.05),
                            states, 3.0, T, R, rng=np.random.default_rng(9))
    assert c.shape == (S, R, T, 2) and x.shape == (S, R, T, n)
    seed = engine.draw_seed(np.random.default_rng(9))
    gen = _RingGenerator(n, seed=5)
    t_eval = np.linspace(0, 3.0, T)
    for pair in range(S * R):
        j, i = divmod(pair, R)
        nl = calculate_neighbor_list(gen())
        p = CNVMParameters(num_opinions=2, network=nl, r=1.0, r_tilde=0.05)
        model = CNVM(p)
        rate, noise_p = model.event_rates()
        rp, col = oracle.neighbor_list_to_csr(nl)
        nbtab, nblg = oracle.build_neighbor_table(rp, col)
        counts = oracle.native_event_counts(seed, pair, 1, t_eval, rate)[0]
        ox, oc, _, _ = oracle.cnvm_native(states[j], 2, rp, col, int(oracle.prob_to_thr(noise_p)), oracle.prob_to_thr(p.prob_imit),
                                          oracle.prob_to_thr(p.prob_noise), None, None, counts, seed, pair, nbtab=nbtab,
                                          nbtab_log2=max(nblg, 0))
        np.testing.assert_array_equal(x[j, i], ox)
        np.testing.assert_array_equal(c[j, i], oc.astype(float))


def test_update_rates_and_cntm_drivers():
    import networkx as nx

    from sponet_b200 import (CNTMParameters, CNVM, CNVMParameters, OpinionShares, estimate_steady_state,
                             sample_many_runs, sample_moments)

    params = CNVMParameters(num_opinions=2, num_agents=80, r=1.0, r_tilde=0.0)
    model = CNVM(params)
    model.update_rates(r_tilde=0.5)
    assert params.r_noise == 1.0 and params.r_imit == 1.0
    x0 = np.zeros(80, dtype=int)
    _, x = model.simulate(20.0, x0, 3, rng=np.random.default_rng(0))
    assert x[-1].sum() > 0  # noise now creates opinion 1 out of consensus
    p = CNTMParameters(nx.erdos_renyi_graph(90, 0.1, seed=2), 1.0, 0.3, 0.5, 0.5)
    s0 = np.random.default_rng(1).integers(0, 2, 90)
    _, a = sample_many_runs(p, s0, 2.0, 4, 3, n_jobs=2, rng=np.random.default_rng(4), mode="replay")
    assert a.shape == (3, 4, 90) and a.dtype == np.uint8
    t, mean, var, n = sample_moments(p, s0, 2.0, 4, 200, rel_tol=0.2, collective_variable=OpinionShares(2, True),
                                     rng=np.random.default_rng(5))
    assert mean.shape == (4, 2) and np.allclose(mean.sum(1), 1.0) and n >= 200
    est = estimate_steady_state(p, OpinionShares(2, normalize=True), tol_abs=0.1, rng=np.random.default_rng(6), verbose=False)
    assert est.shape == (2,) and abs(est.sum() - 1) < 1e-12


def test_grouped_and_weighted_cvs_are_evaluated_in_kernel():
    """DegreeWeightedOpinionShares, OpinionSharesByDegree, integer-weighted OpinionShares and composites
    over one table are kept by the simulation kernel (integer tables, exact); the result must equal the
    CV evaluated on the stored states of the same chains, bit for bit -- also with a ring of warps."""
    import networkx as nx

    from sponet_b200 import (CNVM, CNVMParameters, CompositeCollectiveVariable, DegreeWeightedOpinionShares,
                             OpinionShares, OpinionSharesByDegree, engine, sample_many_runs, sample_moments)
    from sponet_b200.multiprocessing import _kernel_cv_spec

    g = nx.barabasi_albert_graph(300, 3, seed=2)
    params = CNVMParameters(num_opinions=3, network=g, r=R3, r_tilde=RT3, alpha=0.5)
    x0 = np.random.default_rng(0).integers(0, 3, 300)
    wi = np.random.default_rng(1).integers(-3, 8, 300)
    cvs = [
        DegreeWeightedOpinionShares(3, g),
        DegreeWeightedOpinionShares(3, g, normalize=True, idx_to_return=[2, 0]),
        OpinionSharesByDegree(3, g),
        OpinionSharesByDegree(3, g, normalize=True, idx_to_return=1),
        OpinionShares(3, weights=wi, normalize=True),
        CompositeCollectiveVariable([OpinionShares(3), OpinionShares(3, normalize=True, idx_to_return=1)]),
        CompositeCollectiveVariable([OpinionSharesByDegree(3, g), OpinionSharesByDegree(3, g, True, [0])]),
    ]
    _, xs = sample_many_runs(params, x0, 2.0, 4, 6, rng=np.random.default_rng(9))
    for cv in cvs:
        assert _kernel_cv_spec(cv, 300) is not None, type(cv).__name__
        _, c = sample_many_runs(params, x0, 2.0, 4, 6, collective_variable=cv, rng=np.random.default_rng(9))
        assert c.shape == (6, 4, cv.dimension)
        np.testing.assert_array_equal(c, cv(xs.reshape(-1, 300)).reshape(c.shape))
    # float weights and mixed tables are NOT in-kernel (evaluated from device-resident slabs instead)
    assert _kernel_cv_spec(OpinionShares(3, weights=np.random.default_rng(2).random(300)), 300) is None
    assert _kernel_cv_spec(CompositeCollectiveVariable([OpinionShares(3), DegreeWeightedOpinionShares(3, g)]), 300) is None
    # moments of a grouped CV: fused integer sums == moments of the per-run values
    cv = OpinionSharesByDegree(3, g, normalize=True)
    _, c = sample_many_runs(params, x0, 2.0, 4, 300, collective_variable=cv, rng=np.random.default_rng(4))
    _, m, v, n = sample_moments(params, x0, 2.0, 4, 300, rel_tol=1e9, collective_variable=cv, rng=np.random.default_rng(4))
    np.testing.assert_allclose(m, c.mean(0), rtol=1e-12)
    np.testing.assert_allclose(v, c.var(0), rtol=1e-8, atol=1e-14)
    # ring of warps + grouped table (larger chain so that the ring is used)
    n_big = 40000
    gb = nx.fast_gnp_random_graph(n_big, 8 / n_big, seed=1)
    for u in list(nx.isolates(gb)):
        gb.add_edge(u, (u + 1) % n_big)
    pb = CNVMParameters(num_opinions=3, network=gb, r=R3, r_tilde=RT3)
    xb = np.random.default_rng(5).integers(0, 3, n_big).astype(np.uint8)
    spec = _kernel_cv_spec(OpinionSharesByDegree(3, gb), n_big)
    mb = CNVM(pb)
    struct, keep = mb._struct()
    outs = [engine.run_native("cnvm", mb._graph(), n_big, struct, xb[None], np.linspace(0, 0.05, 3), 3, 40, 0, 40,
                              cv_spec=spec, want_cv=True, warps_per_chain=w)[1] for w in (1, 4)]
    np.testing.assert_array_equal(outs[0], outs[1])
    del keep
