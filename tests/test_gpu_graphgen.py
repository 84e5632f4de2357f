"""GPU: per-run networks generated on the device (reference sponet/network_generator.py:26-85 called
per run through cnvm/model.py:95-96 by multiprocessing.py:173-178).

Bars: every graph of a device batch equals the sequential statement (oracle_er_graph) bit for bit; chains launched on a
batch (one graph per chain) equal the same chains launched one by one on the extracted graphs -- states, CV and counters
bit for bit -- and the sequential chain oracle; `sample_many_runs` with a device-capable generator does not depend on
the block size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,force,S,R,rb,re", [
    (300, 0.012, True, 2, 9, 2, 7),     # mean degree 3.6: ~8 isolates per graph, repaired
    (1000, 0.01, True, 1, 5, 0, 5),
    (300, 0.03, False, 3, 4, 1, 4),     # rejection: ~4 % of the samples have an isolate
    (64, 1.0, False, 1, 2, 0, 2),       # complete graphs
    (5000, 0.002, True, 1, 3, 0, 3),
    (400, 0.3, False, 1, 2, 0, 2),      # dense
])
def test_device_batch_equals_oracle(oracle, n, p, force, S, R, rb, re):
    from sponet_b200 import engine

    seed = 1234567
    b = engine.GraphBatch(n, p, seed, S, R, rb, re, force_no_isolates=force, padded_rows=True)
    assert b.count == S * (re - rb)
    dmin, dmax = 1 << 30, 0
    for slot in range(b.count):
        j, i = divmod(slot, re - rb)
        rp, col = b.csr(slot)
        orp, ocol, _ = oracle.er_graph(n, p, seed, j * R + rb + i, force_no_isolates=force)
        np.testing.assert_array_equal(rp, orp)
        np.testing.assert_array_equal(col, ocol)
        dmin, dmax = min(dmin, int(np.diff(rp).min())), max(dmax, int(np.diff(rp).max()))
    info = b.info()
    assert info["min_degree"] == dmin >= 1 and info["max_degree"] == dmax and info["num_agents"] == n
    # a graph is a function of (seed, index): the same index inside another batch
    b2 = engine.GraphBatch(n, p, seed, 1, S * R, 0, S * R, force_no_isolates=force)
    for slot in (0, b.count - 1):
        j, i = divmod(slot, re - rb)
        for x, y in zip(b.csr(slot), b2.csr(j * R + rb + i)):
            np.testing.assert_array_equal(x, y)
    b.close()
    b2.close()


def test_device_batch_gives_up_like_the_reference():
    from sponet_b200 import engine

    with pytest.raises(RuntimeError, match="Timeout"):
        engine.GraphBatch(400, 0.008, 1, 1, 4, 0, 4, force_no_isolates=False, max_attempts=3)
    with pytest.raises(ValueError):
        engine.GraphBatch(400, 0.0, 1, 1, 4, 0, 4)


def _single(kind, model_kw, rp, col, x0, t_eval, seed, R_total, c, want_x=True, cv_spec=None):
    """Chain id c alone on its own graph through the ordinary single-graph path."""
    from conftest import csr_to_neighbor_list
    from sponet_b200 import CNTM, CNTMParameters, CNVM, CNVMParameters, engine

    nl = csr_to_neighbor_list(rp, col)
    if kind == "cnvm":
        model = CNVM(CNVMParameters(network=nl, **model_kw))
        struct, keep = model._struct()
    else:
        import networkx as nx

        g = nx.Graph()
        g.add_nodes_from(range(len(nl)))
        g.add_edges_from((i, int(j)) for i, nb in enumerate(nl) for j in nb if i < j)
        model = CNTM(CNTMParameters(network=g, **model_kw))
        struct, keep = model._struct(), None
    out = engine.run_native(kind, model._graph(), len(nl), struct, x0[None, :], t_eval, seed, R_total, c, c + 1,
                            want_x=want_x, want_cv=cv_spec is not None, cv_spec=cv_spec)
    del keep
    return out, model


def _oracle_chain(oracle, kw, rp, col, x0, t_eval, seed, c):
    """Chain id c on the CSR (rp, col) by the sequential oracle (neighbour = col[begin + mulhi(r2, degree)], the rule the
    kernels follow on a graph without sampler table -- a batch never has one)."""
    from conftest import csr_to_neighbor_list
    from sponet_b200 import CNVM, CNVMParameters

    model = CNVM(CNVMParameters(network=csr_to_neighbor_list(rp, col), **kw))
    rate, noise_p = model.event_rates()
    counts = oracle.native_event_counts(seed, c, 1, t_eval, rate)[0]
    prm = model.params
    da = np.asarray(model.degree_alpha, dtype=float)
    alias_thr = alias = None
    if not np.all(da == da[0]):  # alpha != 1: the graph's own alias table (sampling.py:54-96)
        prob, alias = oracle.build_alias_table(da)
        alias_thr = oracle.prob_to_thr(prob)
    ox, oc, _, _ = oracle.cnvm_native(np.ascontiguousarray(x0, dtype=np.uint8), prm.num_opinions, rp, col, int(oracle.prob_to_thr(noise_p)),
                                      oracle.prob_to_thr(prm.prob_imit), oracle.prob_to_thr(prm.prob_noise), alias_thr, alias, counts,
                                      seed, c)
    return ox, oc


@pytest.mark.parametrize("records", [True, False])
@pytest.mark.parametrize("M,n,p", [(2, 600, 0.01), (3, 2000, 0.004), (5, 400, 0.03), (2, 300, 0.4), (3, 700, 0.5), (2, 3000, 0.006)])
def test_cnvm_chains_on_a_batch_equal_single_graph_chains(oracle, M, n, p, records):
    """(records: the 64-byte agent records, neighbours 0..14 inside, the rest out of the row -- mean degrees 6 .. 350 here
    cover both sides and the fallback without records above degree 255.)"""
    from sponet_b200 import CNVM, CNVMParameters, OpinionShares, engine
    from sponet_b200.multiprocessing import _kernel_cv_spec

    rng = np.random.default_rng(M)
    S, R, rb, re, T = 2, 7, 1, 6, 5
    r = rng.uniform(0.5, 2.0, (M, M))
    rt = rng.uniform(0.01, 0.1, (M, M))
    kw = dict(num_opinions=M, r=r, r_tilde=rt)
    states = rng.integers(0, M, (S, n)).astype(np.uint8)
    t_eval = np.linspace(0.0, 4.0, T)
    seed, gseed = 4242, 777
    batch = engine.GraphBatch(n, p, gseed, S, R, rb, re, force_no_isolates=True, records=records)
    model = CNVM(CNVMParameters(num_agents=n, **kw))  # alpha = 1: uniform agent weights whatever the graph
    struct, keep = model._struct()
    spec = _kernel_cv_spec(OpinionShares(M), n)
    counters = np.zeros(8, dtype=np.uint64)
    x, cv = engine.run_native("cnvm", batch, n, struct, states, t_eval, seed, R, rb, re, want_x=True, want_cv=True, cv_spec=spec,
                              counters=counters)
    assert x.shape == (S, re - rb, T, n) and cv.shape == (S, re - rb, T, M)
    for j in range(S):
        for i in range(re - rb):
            rp, col = batch.csr(j * (re - rb) + i)
            c = j * R + rb + i
            ox, oc = _oracle_chain(oracle, kw, rp, col, states[j], t_eval, seed, c)
            np.testing.assert_array_equal(x[j, i], ox)
            np.testing.assert_array_equal(cv[j, i], oc.astype(float))
            if np.diff(rp).max() > 64:  # no sampler table on the single graph either: the ordinary path, same chain
                (sx, scv), _ = _single("cnvm", kw, rp, col, states[j], t_eval, seed, S * R, c, cv_spec=spec)
                np.testing.assert_array_equal(x[j, i], sx[0, 0])
                np.testing.assert_array_equal(cv[j, i], scv[0, 0])
    assert counters[0] > 0
    batch.close()
    del keep
    # a batch of the wrong size is refused
    small = engine.GraphBatch(n, p, gseed, 1, R, 0, 2, force_no_isolates=True)
    with pytest.raises(ValueError, match="graph batch holds"):
        engine.run_native("cnvm", small, n, struct, states, t_eval, seed, R, rb, re, want_x=True)
    small.close()


def test_cntm_chains_on_a_batch_equal_single_graph_chains():
    from sponet_b200 import engine

    rng = np.random.default_rng(5)
    n, p, S, R, T = 500, 0.02, 2, 4, 4
    kw = dict(r=1.0, r_tilde=0.2, threshold_01=0.5, threshold_10=0.4)
    states = rng.integers(0, 2, (S, n)).astype(np.uint8)
    t_eval = np.linspace(0.0, 3.0, T)
    batch = engine.GraphBatch(n, p, 99, S, R, 0, R, force_no_isolates=True, padded_rows=True)
    struct = engine.cntm_struct(1 / (n * 1.2), 0.2 / 1.2, type("P", (), kw))
    x, _ = engine.run_native("cntm", batch, n, struct, states, t_eval, 31, R, 0, R, want_x=True)
    for j in range(S):
        for i in range(R):
            rp, col = batch.csr(j * R + i)
            (sx, _), _ = _single("cntm", kw, rp, col, states[j], t_eval, 31, S * R, j * R + i)
            np.testing.assert_array_equal(x[j, i], sx[0, 0])
    nopad = engine.GraphBatch(n, p, 99, S, R, 0, R, force_no_isolates=True)
    with pytest.raises(ValueError, match="padded rows"):
        engine.run_native("cntm", nopad, n, struct, states, t_eval, 31, R, 0, R, want_x=True)
    batch.close()
    nopad.close()


def test_sample_many_runs_with_a_device_generator(oracle):
    """The public path: ErdosRenyiGenerator in the parameters -> device batches, one launch per block; equal to the chains
    computed one by one on the graphs of the same batch seed; independent of the block size; rng advanced once."""
    from sponet_b200 import CNVMParameters, OpinionShares, engine, sample_many_runs, sample_moments
    from sponet_b200 import multiprocessing as mp
    from sponet_b200.network_generator import ErdosRenyiGenerator

    n, p, S, R, T, M = 400, 0.02, 2, 6, 4, 3
    kw = dict(num_opinions=M, r=[[0, 1, 2], [1, 0, 1], [2, 0, 0]], r_tilde=[[0, .2, .1], [0, 0, .1], [.1, .2, 0]])

    def params(seed):
        return CNVMParameters(network_generator=ErdosRenyiGenerator(n, p, rng=np.random.default_rng(seed), force_no_isolates=True), **kw)

    states = np.random.default_rng(0).integers(0, M, (S, n))
    cv = OpinionShares(M)
    t, c = sample_many_runs(params(8), states, 2.0, T, R, collective_variable=cv, rng=np.random.default_rng(1))
    _, x = sample_many_runs(params(8), states, 2.0, T, R, rng=np.random.default_rng(1))
    assert c.shape == (S, R, T, M) and x.shape == (S, R, T, n) and x.dtype == np.uint8
    for j in range(S):
        for i in range(R):
            np.testing.assert_array_equal(c[j, i], np.stack([np.bincount(s, minlength=M) for s in x[j, i]]).astype(float))
    # chain (j, i) on graph index j R + i of the batch seed the generator's rng yields first
    gseed = ErdosRenyiGenerator(n, p, rng=np.random.default_rng(8)).draw_batch_seed()
    seed = engine.draw_seed(np.random.default_rng(1))
    whole = engine.GraphBatch(n, p, gseed, S, R, 0, R, force_no_isolates=True)
    for j, i in ((0, 0), (1, 2), (1, R - 1)):
        rp, col = whole.csr(j * R + i)
        ox, _ = _oracle_chain(oracle, kw, rp, col, states[j], t, seed, j * R + i)
        np.testing.assert_array_equal(x[j, i], ox)
    whole.close()
    # small blocks (1 state x 2 runs per launch)
    old = mp._BATCH_BYTES
    try:
        mp._BATCH_BYTES = int(2.5 * (84.0 * n + 4.0 * n * (p * (n - 1) + 2.0)))
        _, c2 = sample_many_runs(params(8), states, 2.0, T, R, collective_variable=cv, rng=np.random.default_rng(1))
        mp._BATCH_BYTES = int(1.5 * (84.0 * n + 4.0 * n * (p * (n - 1) + 2.0)))  # one graph per launch: blocks of states too
        _, c3 = sample_many_runs(params(8), states, 2.0, T, R, collective_variable=cv, rng=np.random.default_rng(1))
    finally:
        mp._BATCH_BYTES = old
    np.testing.assert_array_equal(c, c2)
    np.testing.assert_array_equal(c, c3)
    # a 1-D initial state is squeezed; sample_moments rides on the same path
    _, c1 = sample_many_runs(params(8), states[0], 2.0, T, R, collective_variable=cv, rng=np.random.default_rng(1))
    np.testing.assert_array_equal(c1, c[0])
    _, mean, var, ns = sample_moments(params(8), states[0], 2.0, T, R, rel_tol=1e9, collective_variable=cv, rng=np.random.default_rng(1))
    assert ns == R
    np.testing.assert_allclose(mean, c[0].mean(0), atol=1e-12)
    # a collective variable the kernels do not evaluate (its own fixed graph): from the returned states of the same chains
    import networkx as nx

    from sponet_b200 import Interfaces

    kw2 = dict(num_opinions=2, r=[[0, 1.0], [1.2, 0]], r_tilde=[[0, 0.05], [0.02, 0]])
    iface = Interfaces(nx.cycle_graph(n))

    def params2(seed):
        return CNVMParameters(network_generator=ErdosRenyiGenerator(n, p, rng=np.random.default_rng(seed), force_no_isolates=True), **kw2)

    s2 = states % 2
    _, ci = sample_many_runs(params2(8), s2, 2.0, T, 3, collective_variable=iface, rng=np.random.default_rng(1))
    _, xi = sample_many_runs(params2(8), s2, 2.0, T, 3, rng=np.random.default_rng(1))
    assert ci.shape == (S, 3, T, 1)
    np.testing.assert_array_equal(ci[1, 2], iface(xi[1, 2]))
    # alpha != 1: agent weights, event rate and noise probability of every chain come from its own graph
    for alpha in (0.5, 0.0, 2.0):
        pa = CNVMParameters(network_generator=ErdosRenyiGenerator(n, p, rng=np.random.default_rng(8), force_no_isolates=True), alpha=alpha, **kw)
        ta, xa = sample_many_runs(pa, states, 1.5, 3, 3, rng=np.random.default_rng(1))
        pa = CNVMParameters(network_generator=ErdosRenyiGenerator(n, p, rng=np.random.default_rng(8), force_no_isolates=True), alpha=alpha, **kw)
        _, ca = sample_many_runs(pa, states, 1.5, 3, 3, collective_variable=cv, rng=np.random.default_rng(1))
        assert xa.shape == (S, 3, 3, n) and ca.shape == (S, 3, 3, M)
        wb = engine.GraphBatch(n, p, gseed, S, 3, 0, 3, force_no_isolates=True)
        for j, i in ((0, 0), (1, 2)):
            rp, col = wb.csr(j * 3 + i)
            ox, oc = _oracle_chain(oracle, dict(alpha=alpha, **kw), rp, col, states[j], ta, seed, j * 3 + i)
            np.testing.assert_array_equal(xa[j, i], ox)
            np.testing.assert_array_equal(ca[j, i], oc.astype(float))
        wb.close()


def test_per_run_networks_agree_with_the_host_pipeline_statistically():
    """Same ensemble through device-generated graphs and through networkx graphs (the reference's generator path):
    the two are different samples of the same process -- means of the final shares within 5 sigma."""
    from sponet_b200 import CNVMParameters, OpinionShares, sample_many_runs
    from sponet_b200 import multiprocessing as mp
    from sponet_b200.network_generator import ErdosRenyiGenerator

    n, p, R, M = 300, 0.03, 400, 2
    kw = dict(num_opinions=M, r=[[0, 1.0], [1.2, 0]], r_tilde=[[0, 0.05], [0.02, 0]])
    x0 = np.random.default_rng(0).integers(0, M, n)
    cv = OpinionShares(M, normalize=True)

    def run(force_host):
        prm = CNVMParameters(network_generator=ErdosRenyiGenerator(n, p, rng=np.random.default_rng(3), force_no_isolates=True), **kw)
        mp._FORCE_HOST_NETWORKS = force_host
        try:
            return sample_many_runs(prm, x0, 6.0, 4, R, collective_variable=cv, rng=np.random.default_rng(2))[1]
        finally:
            mp._FORCE_HOST_NETWORKS = False

    a, b = run(False), run(True)
    for k in range(1, 4):
        da, db = a[:, k, 0], b[:, k, 0]
        z = (da.mean() - db.mean()) / np.sqrt(da.var() / R + db.var() / R)
        assert abs(z) < 5.0, (k, z)
        assert 0.6 < da.var() / db.var() < 1.6


def test_h_sized_batch_beyond_2_31_entries(oracle):
    """2 368 graphs of H's size in one batch: 2.37e9 adjacency entries, i.e. row offsets beyond 2^31 (they are unsigned 32-bit)
    and 15 GB of records.  The last graphs equal the sequential generator, and chains on them the sequential chain oracle."""
    from sponet_b200 import CNVM, CNVMParameters, engine

    n, G, M = 100_000, 2368, 3
    p = 10 / (n - 1)
    kw = dict(num_opinions=M, r=[[0, 1, 2], [1, 0, 1], [2, 0, 0]], r_tilde=[[0, .2, .1], [0, 0, .1], [.1, .2, 0]])
    batch = engine.GraphBatch(n, p, 31337, 1, G, 0, G, force_no_isolates=True)
    assert batch.info()["nnz"] > 2**31
    x0 = np.random.default_rng(0).integers(0, M, n).astype(np.uint8)
    t_eval = np.linspace(0.0, 1.0, 3)
    model = CNVM(CNVMParameters(num_agents=n, **kw))
    struct, keep = model._struct()
    x, _ = engine.run_native("cnvm", batch, n, struct, x0[None, :], t_eval, 5, G, 0, G, want_x=True, final_state_only=True)
    assert x.shape[:2] == (1, G)
    for slot in (0, 2200, G - 1):
        rp, col = batch.csr(slot)
        orp, ocol, _ = oracle.er_graph(n, p, 31337, slot, force_no_isolates=True)
        np.testing.assert_array_equal(rp, orp)
        np.testing.assert_array_equal(col, ocol)
        ox, _ = _oracle_chain(oracle, kw, rp, col, x0, t_eval, 5, slot)
        np.testing.assert_array_equal(x[0, slot].reshape(-1, n)[-1], ox[-1])
    batch.close()
    del keep
    engine.trim_graph_batch_memory()


@pytest.mark.parametrize("n,m,S,R,rb,re", [(500, 3, 2, 5, 1, 4), (4000, 5, 1, 3, 0, 3), (300, 1, 1, 2, 0, 2), (200, 64, 1, 2, 0, 2)])
def test_ba_device_batch_equals_oracle(oracle, n, m, S, R, rb, re):
    from sponet_b200 import engine

    b = engine.GraphBatch.barabasi_albert(n, m, 4711, S, R, rb, re, padded_rows=True)
    for slot in range(b.count):
        j, i = divmod(slot, re - rb)
        rp, col = b.csr(slot)
        orp, ocol = oracle.ba_graph(n, m, 4711, j * R + rb + i)
        np.testing.assert_array_equal(rp, orp)
        np.testing.assert_array_equal(col, ocol)
    info = b.info()
    assert info["min_degree"] >= 1 and info["nnz"] == b.count * 2 * m * (n - m)
    b.close()
    with pytest.raises(ValueError, match="m >= 1 and m < n"):
        engine.GraphBatch.barabasi_albert(10, 10, 1, 1, 1, 0, 1)


@pytest.mark.parametrize("records", [True, False])
def test_cnvm_chains_on_a_ba_batch(oracle, records):
    """Scale-free per-run graphs: hubs of several hundred neighbours (record header with more than 8 degree bits, most of a
    hub's neighbours outside its record)."""
    from sponet_b200 import CNVM, CNVMParameters, OpinionShares, engine
    from sponet_b200.multiprocessing import _kernel_cv_spec

    n, m, M, S, R, T = 5000, 4, 3, 2, 4, 4
    kw = dict(num_opinions=M, r=[[0, 1, 2], [1, 0, 1], [2, 0, 0]], r_tilde=[[0, .2, .1], [0, 0, .1], [.1, .2, 0]])
    states = np.random.default_rng(2).integers(0, M, (S, n)).astype(np.uint8)
    t_eval = np.linspace(0.0, 3.0, T)
    batch = engine.GraphBatch.barabasi_albert(n, m, 5, S, R, 0, R, records=records)
    assert batch.info()["max_degree"] > 255
    model = CNVM(CNVMParameters(num_agents=n, **kw))
    struct, keep = model._struct()
    spec = _kernel_cv_spec(OpinionShares(M), n)
    x, cv = engine.run_native("cnvm", batch, n, struct, states, t_eval, 77, R, 0, R, want_x=True, want_cv=True, cv_spec=spec)
    for j in range(S):
        for i in range(R):
            rp, col = batch.csr(j * R + i)
            ox, oc = _oracle_chain(oracle, kw, rp, col, states[j], t_eval, 77, j * R + i)
            np.testing.assert_array_equal(x[j, i], ox)
            np.testing.assert_array_equal(cv[j, i], oc.astype(float))
    batch.close()
    del keep


def test_sample_many_runs_with_a_ba_generator(oracle):
    from sponet_b200 import BarabasiAlbertGenerator, CNVMParameters, OpinionShares, engine, sample_many_runs

    n, m, R, T, M = 800, 3, 5, 3, 2
    kw = dict(num_opinions=M, r=[[0, 1.0], [1.2, 0]], r_tilde=[[0, 0.05], [0.02, 0]])
    x0 = np.random.default_rng(0).integers(0, M, n)

    def params():
        return CNVMParameters(network_generator=BarabasiAlbertGenerator(n, m, rng=np.random.default_rng(6)), **kw)

    t, c = sample_many_runs(params(), x0, 2.0, T, R, collective_variable=OpinionShares(M), rng=np.random.default_rng(1))
    _, x = sample_many_runs(params(), x0, 2.0, T, R, rng=np.random.default_rng(1))
    assert c.shape == (R, T, M) and x.shape == (R, T, n)
    gseed = BarabasiAlbertGenerator(n, m, rng=np.random.default_rng(6)).draw_batch_seed()
    seed = engine.draw_seed(np.random.default_rng(1))
    for i in (0, R - 1):
        rp, col = oracle.ba_graph(n, m, gseed, i)
        ox, oc = _oracle_chain(oracle, kw, rp, col, x0, t, seed, i)
        np.testing.assert_array_equal(x[i], ox)
        np.testing.assert_array_equal(c[i], oc.astype(float))


@pytest.mark.parametrize("n,k,p,S,R,rb,re", [(500, 6, 0.2, 2, 5, 1, 4), (3000, 10, 0.1, 1, 3, 0, 3), (200, 4, 1.0, 1, 3, 0, 3),
                                             (64, 2, 0.0, 1, 2, 0, 2), (400, 2, 0.5, 1, 3, 0, 3), (100, 40, 0.5, 1, 2, 0, 2)])
def test_ws_device_batch_equals_oracle(oracle, n, k, p, S, R, rb, re):
    """(400, 2, 0.5): sparse rings fall apart and are drawn again -- the attempt counter; (100, 40, .5): rows of 2 x 32+ lanes.)"""
    from sponet_b200 import engine

    b = engine.GraphBatch.watts_strogatz(n, k, p, 815, S, R, rb, re, padded_rows=True)
    for slot in range(b.count):
        j, i = divmod(slot, re - rb)
        rp, col = b.csr(slot)
        orp, ocol, _ = oracle.ws_graph(n, k, p, 815, j * R + rb + i)
        np.testing.assert_array_equal(rp, orp)
        np.testing.assert_array_equal(col, ocol)
    assert b.info()["nnz"] == b.count * n * (k // 2) * 2
    b.close()
    with pytest.raises(RuntimeError, match="Maximum number of tries"):
        engine.GraphBatch.watts_strogatz(2000, 2, 0.9, 3, 1, 2, 0, 2, tries=3)
    with pytest.raises(ValueError):
        engine.GraphBatch.watts_strogatz(10, 10, 0.1, 3, 1, 1, 0, 1)


def test_sample_many_runs_with_a_ws_generator(oracle):
    from sponet_b200 import CNVMParameters, OpinionShares, WattsStrogatzGenerator, engine, sample_many_runs

    n, k, p, R, T, M = 1000, 8, 0.15, 5, 3, 2
    kw = dict(num_opinions=M, r=[[0, 1.0], [1.2, 0]], r_tilde=[[0, 0.05], [0.02, 0]])
    x0 = np.random.default_rng(0).integers(0, M, n)

    def params():
        return CNVMParameters(network_generator=WattsStrogatzGenerator(n, k, p, rng=np.random.default_rng(6)), **kw)

    t, c = sample_many_runs(params(), x0, 2.0, T, R, collective_variable=OpinionShares(M), rng=np.random.default_rng(1))
    _, x = sample_many_runs(params(), x0, 2.0, T, R, rng=np.random.default_rng(1))
    gseed = WattsStrogatzGenerator(n, k, p, rng=np.random.default_rng(6)).draw_batch_seed()
    seed = engine.draw_seed(np.random.default_rng(1))
    for i in (0, R - 1):
        rp, col, _ = oracle.ws_graph(n, k, p, gseed, i)
        ox, oc = _oracle_chain(oracle, kw, rp, col, x0, t, seed, i)
        np.testing.assert_array_equal(x[i], ox)
        np.testing.assert_array_equal(c[i], oc.astype(float))


@pytest.mark.parametrize("n,d,S,R,rb,re", [(500, 4, 2, 5, 1, 4), (3000, 10, 1, 3, 0, 3), (12, 5, 1, 8, 0, 8), (64, 63, 1, 2, 0, 2),
                                           (50, 1, 1, 2, 0, 2), (100, 40, 1, 2, 0, 2)])
def test_regular_device_batch_equals_oracle(oracle, n, d, S, R, rb, re):
    """((12, 5): most attempts fail and restart; (100, 40): rows longer than a warp.)"""
    from sponet_b200 import engine

    b = engine.GraphBatch.random_regular(n, d, 99, S, R, rb, re, padded_rows=True)
    for slot in range(b.count):
        j, i = divmod(slot, re - rb)
        rp, col = b.csr(slot)
        orp, ocol, _ = oracle.regular_graph(n, d, 99, j * R + rb + i)
        np.testing.assert_array_equal(rp, orp)
        np.testing.assert_array_equal(col, ocol)
    info = b.info()
    assert info["min_degree"] == info["max_degree"] == d
    b.close()
    with pytest.raises(ValueError, match="n \\* d must be even"):
        engine.GraphBatch.random_regular(11, 3, 1, 1, 1, 0, 1)


def test_sample_many_runs_with_a_regular_generator(oracle):
    from sponet_b200 import CNVMParameters, OpinionShares, RandomRegularGenerator, engine, sample_many_runs

    n, d, R, T, M = 1000, 6, 4, 3, 2
    kw = dict(num_opinions=M, r=[[0, 1.0], [1.2, 0]], r_tilde=[[0, 0.05], [0.02, 0]])
    x0 = np.random.default_rng(0).integers(0, M, n)

    def params():
        return CNVMParameters(network_generator=RandomRegularGenerator(n, d, rng=np.random.default_rng(6)), **kw)

    t, c = sample_many_runs(params(), x0, 2.0, T, R, collective_variable=OpinionShares(M), rng=np.random.default_rng(1))
    _, x = sample_many_runs(params(), x0, 2.0, T, R, rng=np.random.default_rng(1))
    gseed = RandomRegularGenerator(n, d, rng=np.random.default_rng(6)).draw_batch_seed()
    seed = engine.draw_seed(np.random.default_rng(1))
    for i in (0, R - 1):
        rp, col, _ = oracle.regular_graph(n, d, gseed, i)
        ox, oc = _oracle_chain(oracle, kw, rp, col, x0, t, seed, i)
        np.testing.assert_array_equal(x[i], ox)
        np.testing.assert_array_equal(c[i], oc.astype(float))


def test_weighted_agents_on_scale_free_batches(oracle):
    """alpha = 0 on Barabasi-Albert per-run graphs: agents are picked proportionally to their degree (hubs of several hundred
    neighbours), through per-graph alias tables built on the device -- equal to the sequential oracle with the alias table
    of sponet/sampling.py:54-96 on the same graph."""
    from sponet_b200 import BarabasiAlbertGenerator, CNVMParameters, engine, sample_many_runs

    n, m, R, M = 3000, 3, 4, 3
    kw = dict(num_opinions=M, r=[[0, 1, 2], [1, 0, 1], [2, 0, 0]], r_tilde=[[0, .2, .1], [0, 0, .1], [.1, .2, 0]], alpha=0.0)
    x0 = np.random.default_rng(5).integers(0, M, n)
    t, x = sample_many_runs(CNVMParameters(network_generator=BarabasiAlbertGenerator(n, m, rng=np.random.default_rng(2)), **kw),
                            x0, 0.02, 3, R, rng=np.random.default_rng(3))
    gseed = BarabasiAlbertGenerator(n, m, rng=np.random.default_rng(2)).draw_batch_seed()
    seed = engine.draw_seed(np.random.default_rng(3))
    for i in range(R):
        rp, col = oracle.ba_graph(n, m, gseed, i)
        ox, _ = _oracle_chain(oracle, kw, rp, col, x0, t, seed, i)
        np.testing.assert_array_equal(x[i], ox)
