"""GPU: the CUDA replay kernels (through the C ABI) against the reference fixtures and the oracle.

Bit-exact: agent states and float64 event times.
"""
import ctypes as C

import numpy as np
import pytest
from numpy.random import PCG64, Generator, Philox

from conftest import csr_to_neighbor_list, golden_names, load_golden

pytestmark = pytest.mark.gpu

BITGENS = {"PCG64": PCG64, "Philox": Philox}


def _t_eval_arg(g):
    te = g["t_eval"]
    if te.ndim == 0:
        return None if int(te) < 0 else int(te)
    return te.astype(float)


@pytest.mark.parametrize("name", golden_names("cnvm_"))
def test_cnvm_simulate_replay_matches_reference(name):
    from sponet_b200 import CNVM, CNVMParameters

    g = load_golden(name)
    kw = dict(num_opinions=int(g["num_opinions"]), r=g["r"], r_tilde=g["r_tilde"], alpha=float(g["alpha"]))
    if bool(g["complete"]):
        params = CNVMParameters(num_agents=int(g["num_agents"]), **kw)
    else:
        params = CNVMParameters(network=csr_to_neighbor_list(g["row_ptr"], g["col"]), **kw)
    np.testing.assert_array_equal(params.prob_imit, g["prob_imit"])
    np.testing.assert_array_equal(params.prob_noise, g["prob_noise"])
    model = CNVM(params)
    np.testing.assert_array_equal(np.asarray(model.degree_alpha, dtype=float), g["degree_alpha"])

    rng = Generator(BITGENS[str(g["bitgen"])](int(g["seed"])))
    t, x = model.simulate(float(g["t_max"]), g["x_init"], _t_eval_arg(g), rng=rng, mode="replay")
    assert x.dtype == g["x_traj"].dtype
    np.testing.assert_array_equal(t, g["t_traj"])
    np.testing.assert_array_equal(x, g["x_traj"])


@pytest.mark.parametrize("name", golden_names("cntm_"))
def test_cntm_simulate_replay_matches_reference(name):
    import networkx as nx

    from sponet_b200 import CNTM, CNTMParameters

    g = load_golden(name)
    nl = csr_to_neighbor_list(g["row_ptr"], g["col"])
    net = nx.Graph()
    net.add_nodes_from(range(len(nl)))
    # insertion order of edges must reproduce the neighbour order of the fixture's CSR
    params = CNTMParameters(net, float(g["r"]), float(g["r_tilde"]), float(g["threshold_01"]), float(g["threshold_10"]))
    model = CNTM(params)
    model.neighbor_list = nl  # the fixture's exact rows (networkx order is not reconstructible from CSR)
    assert model.next_event_rate == float(g["next_event_rate"]) and model.noise_prob == float(g["noise_prob"])
    rng = Generator(BITGENS[str(g["bitgen"])](int(g["seed"])))
    t, x = model.simulate(float(g["t_max"]), g["x_init"], _t_eval_arg(g), rng=rng, mode="replay")
    np.testing.assert_array_equal(t, g["t_traj"])
    np.testing.assert_array_equal(x, g["x_traj"])


def test_replay_leaves_rng_where_the_reference_would(oracle):
    """After a replay call the caller's generator has advanced by exactly the consumed draws."""
    from sponet_b200 import CNVM, CNVMParameters

    g = load_golden("cnvm_ba100_alpha1_linspace")
    params = CNVMParameters(num_opinions=3, network=csr_to_neighbor_list(g["row_ptr"], g["col"]), r=g["r"], r_tilde=g["r_tilde"])
    rng = Generator(PCG64(int(g["seed"])))
    CNVM(params).simulate(10.0, g["x_init"], 101, rng=rng, mode="replay")
    s = oracle.Stream(PCG64(int(g["seed"])).random_raw(1_000_000))
    oracle.cnvm_teval(g["x_init"], np.linspace(0, 10, 101), 3, g["row_ptr"], g["col"], float(g["r_imit"]),
                      float(g["r_noise"]), g["prob_imit"], g["prob_noise"], g["degree_alpha"], s)
    ref = PCG64(int(g["seed"]))
    ref.random_raw(s.pos)
    assert rng.bit_generator.state == ref.state


@pytest.mark.parametrize("name", golden_names("smr_"))
def test_sample_many_runs_replay_matches_reference(name):
    from sponet_b200 import CNVMParameters, OpinionShares, sample_many_runs

    g = load_golden(name)
    params = CNVMParameters(num_opinions=3, network=csr_to_neighbor_list(g["row_ptr"], g["col"]), r=g["r"], r_tilde=g["r_tilde"])
    n_jobs = int(g["n_jobs"]) or None
    cv = OpinionShares(3, normalize=True) if g["x_out"].shape[-1] == 3 else None
    t, x = sample_many_runs(params, g["x_init"], float(g["t_max"]), int(g["num_t"]), int(g["num_runs"]), n_jobs=n_jobs,
                            collective_variable=cv, rng=Generator(PCG64(int(g["seed"]))), mode="replay")
    np.testing.assert_array_equal(t, g["t_out"])
    assert x.dtype == g["x_out"].dtype and x.shape == g["x_out"].shape
    np.testing.assert_array_equal(x, g["x_out"])


def test_replay_larger_network_against_oracle(oracle):
    """N = 5000 ER, ~1.3e5 iterations: CUDA replay == oracle (no fixture needed at this size)."""
    import networkx as nx

    from sponet_b200 import CNVM, CNVMParameters

    net = nx.erdos_renyi_graph(5000, 10 / 4999, seed=5)
    for v in list(nx.isolates(net)):
        net.add_edge(v, (v + 1) % 5000)
    r = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
    rt = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)
    for alpha in (1.0, 0.3):
        params = CNVMParameters(num_opinions=3, network=net, r=r, r_tilde=rt, alpha=alpha)
        model = CNVM(params)
        x0 = np.random.default_rng(1).integers(0, 3, 5000)
        t, x = model.simulate(10.0, x0, 21, rng=Generator(PCG64(77)), mode="replay")
        rp, col = oracle.neighbor_list_to_csr(params.network)
        s = oracle.Stream(PCG64(77).random_raw(6_000_000))
        t_o, x_o, _ = oracle.cnvm_teval(x0, np.linspace(0, 10, 21), 3, rp, col, params.r_imit, params.r_noise,
                                         params.prob_imit, params.prob_noise, model.degree_alpha, s)
        np.testing.assert_array_equal(t, t_o)
        np.testing.assert_array_equal(x, x_o)


def test_log1p_device_matches_host_libm():
    """The ziggurat tail uses log1p; over a long stream every tail draw must match the oracle (libm)."""
    import oracle as orc

    from sponet_b200 import CNVM, CNVMParameters

    # long horizon: 50 * 2 * 8000 = 8e5 exponentials -> ~360 tail draws
    params = CNVMParameters(num_opinions=2, num_agents=50, r=1.0, r_tilde=0.5)
    model = CNVM(params)
    x0 = np.zeros(50, dtype=np.uint8)
    T = 2001
    t, x = model.simulate(8000.0, x0, T, rng=Generator(PCG64(5)), mode="replay")
    s = orc.Stream(PCG64(5).random_raw(6_000_000))
    t_o, x_o, it = orc.cnvm_complete_teval(x0, np.linspace(0, 8000.0, T), 2, params.r_imit, params.r_noise,
                                            params.prob_imit, params.prob_noise, model.degree_alpha[0], s)
    assert it > 500_000
    np.testing.assert_array_equal(t, t_o)
    np.testing.assert_array_equal(x, x_o)
