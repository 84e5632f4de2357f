"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (the reference lives at /root/reference and does not travel):

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py

Every fixture stores the inputs of one reference call (network as CSR, parameters, x_init, t_eval,
BitGenerator name + seed) together with what the reference returned (t_traj, x_traj).  The tests
re-create the raw uint64 stream with ``BitGenerator(seed).random_raw`` -- which is exactly what the
reference consumed -- and replay it through the oracle (CPU tests) and the CUDA replay kernels (GPU
tests); both must reproduce t_traj and x_traj bit for bit.
"""
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")

import networkx as nx  # noqa: E402
import numpy as np  # noqa: E402
from numpy.random import PCG64, Generator, Philox  # noqa: E402

import sponet  # noqa: E402  (the reference)
from sponet import CNTM, CNVM, CNTMParameters, CNVMParameters, OpinionShares, sample_many_runs  # noqa: E402
from sponet.collective_variables import DegreeWeightedOpinionShares, OpinionSharesByDegree  # noqa: E402
from sponet.utils import calculate_neighbor_list  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
BITGENS = {"PCG64": PCG64, "Philox": Philox}

# the reference's standard 3-opinion fixture (tests/tests_cnvm/test_cnvm_simulate.py:22-29)
R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)


def csr(neighbor_list):
    deg = np.array([len(nb) for nb in neighbor_list])
    row_ptr = np.zeros(len(neighbor_list) + 1, dtype=np.int32)
    np.cumsum(deg, out=row_ptr[1:])
    col = np.concatenate([np.asarray(nb, dtype=np.int32) for nb in neighbor_list])
    return row_ptr, col


def connected(g):
    g = g.copy()
    iso = list(nx.isolates(g))
    for v in iso:
        g.add_edge(v, (v + 1) % g.number_of_nodes())
    return g


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in kw.items() if k in ("t_traj", "x_traj", "x_out")})


def cnvm_case(name, params, x0, t_max, t_eval, bitgen, seed):
    model = CNVM(params)
    t_traj, x_traj = model.simulate(t_max, x0, t_eval, rng=Generator(BITGENS[bitgen](seed)))
    kw = dict(
        model="cnvm", num_opinions=params.num_opinions, num_agents=params.num_agents, alpha=params.alpha,
        r_imit=params.r_imit, r_noise=params.r_noise, prob_imit=params.prob_imit, prob_noise=params.prob_noise,
        r=params.r, r_tilde=params.r_tilde, degree_alpha=model.degree_alpha, x_init=np.asarray(x0),
        t_max=t_max, t_eval=np.array(-1) if t_eval is None else np.asarray(t_eval), bitgen=bitgen, seed=seed,
        t_traj=t_traj, x_traj=x_traj, complete=params.network is None,
    )
    if params.network is not None:
        kw["row_ptr"], kw["col"] = csr(params.network)
    save(name, **kw)


def cntm_case(name, g, r, r_tilde, th01, th10, x0, t_max, t_eval, bitgen, seed):
    params = CNTMParameters(network=g, r=r, r_tilde=r_tilde, threshold_01=th01, threshold_10=th10)
    model = CNTM(params)
    t_traj, x_traj = model.simulate(t_max, x0, t_eval, rng=Generator(BITGENS[bitgen](seed)))
    row_ptr, col = csr(calculate_neighbor_list(g))
    save(
        name, model="cntm", num_agents=params.num_agents, r=r, r_tilde=r_tilde, threshold_01=th01,
        threshold_10=th10, next_event_rate=model.next_event_rate, noise_prob=model.noise_prob,
        row_ptr=row_ptr, col=col, x_init=np.asarray(x0), t_max=t_max,
        t_eval=np.array(-1) if t_eval is None else np.asarray(t_eval), bitgen=bitgen, seed=seed,
        t_traj=t_traj, x_traj=x_traj,
    )


def main():
    # ---- CNVM on networks --------------------------------------------------------------------
    g = connected(nx.barabasi_albert_graph(100, 3, seed=1))
    p = CNVMParameters(num_opinions=3, network=g, r=R3, r_tilde=RT3, alpha=1.0)
    x0 = np.random.default_rng(0).integers(0, 3, 100)
    cnvm_case("cnvm_ba100_alpha1_linspace", p, x0, 10.0, 101, "PCG64", 1)
    cnvm_case("cnvm_ba100_alpha1_teval", p, x0, 10.0, np.array([0.5, 1.0, 3.0, 7.5, 10.0]), "Philox", 2)
    cnvm_case("cnvm_ba100_alpha1_all", p, x0, 2.0, None, "PCG64", 3)
    # fine grid: exercises the at-most-one-index-per-iteration drift
    cnvm_case("cnvm_ba100_alpha1_finegrid", p, x0, 1.0, 501, "PCG64", 4)

    g = connected(nx.erdos_renyi_graph(150, 0.06, seed=2))
    p = CNVMParameters(num_opinions=3, network=g, r=R3, r_tilde=RT3, alpha=0.5)
    x0 = np.random.default_rng(1).integers(0, 3, 150)
    cnvm_case("cnvm_er150_alpha05_linspace", p, x0, 20.0, 41, "PCG64", 5)
    p = CNVMParameters(num_opinions=3, network=g, r=R3, r_tilde=RT3, alpha=0.0)
    cnvm_case("cnvm_er150_alpha0_linspace", p, x0, 5.0, 11, "PCG64", 6)

    g = nx.cycle_graph(60)
    p = CNVMParameters(num_opinions=2, network=g, r=1.0, r_tilde=0.01)
    x0 = np.random.default_rng(2).integers(0, 2, 60)
    cnvm_case("cnvm_cycle60_m2", p, x0, 50.0, 51, "PCG64", 7)

    # a larger one so that ziggurat tails and alias branches are all hit (~2.6e5 iterations)
    g = connected(nx.erdos_renyi_graph(2000, 10 / 1999, seed=3))
    p = CNVMParameters(num_opinions=3, network=g, r=R3, r_tilde=RT3, alpha=1.0)
    x0 = np.random.default_rng(3).integers(0, 3, 2000)
    cnvm_case("cnvm_er2000_long", p, x0, 50.0, 6, "PCG64", 8)

    # ---- CNVM complete graph -------------------------------------------------------------------
    p = CNVMParameters(num_opinions=3, num_agents=200, r=R3, r_tilde=RT3)
    x0 = np.random.default_rng(4).integers(0, 3, 200)
    cnvm_case("cnvm_complete200_linspace", p, x0, 20.0, 51, "PCG64", 9)
    cnvm_case("cnvm_complete200_all", p, x0, 1.0, None, "Philox", 10)
    p = CNVMParameters(num_opinions=2, num_agents=1000, r=1.0, r_tilde=0.01)  # BASELINE config 1 rates
    x0 = np.random.default_rng(5).integers(0, 2, 1000)
    cnvm_case("cnvm_complete1000_m2", p, x0, 100.0, 101, "PCG64", 11)
    p = CNVMParameters(num_opinions=257, num_agents=60, r=1.0, r_tilde=0.01)  # uint16 states
    x0 = np.random.default_rng(6).integers(0, 257, 60)
    cnvm_case("cnvm_complete60_m257", p, x0, 5.0, 6, "PCG64", 12)
    # absorbing: r = [[0,1],[0,0]], r_tilde = 0 (tests/tests_cnvm/test_cnvm_simulate.py:181-213)
    p = CNVMParameters(num_opinions=2, num_agents=50, r=np.array([[0, 1], [0, 0]]), r_tilde=0)
    x0 = np.random.default_rng(7).integers(0, 2, 50)
    cnvm_case("cnvm_complete50_absorbing", p, x0, 200.0, 21, "PCG64", 13)

    # ---- CNTM ----------------------------------------------------------------------------------
    g = nx.erdos_renyi_graph(100, 0.1, seed=123)  # tests/tests_cntm/test_cntm_simulate.py:21-29
    x0 = np.random.default_rng(8).integers(0, 2, 100)
    cntm_case("cntm_er100_linspace", g, 1, 0.2, 0.5, 0.5, x0, 50.0, 101, "PCG64", 14)
    cntm_case("cntm_er100_all", g, 1, 0.2, 0.5, 0.5, x0, 5.0, None, "PCG64", 15)
    g = nx.barabasi_albert_graph(300, 5, seed=1)
    x0 = np.random.default_rng(9).integers(0, 2, 300)
    cntm_case("cntm_ba300_thresholds", g, 1, 0.1, 0.3, 0.6, x0, 20.0, 21, "Philox", 16)
    g = nx.erdos_renyi_graph(80, 0.02, seed=5)  # has isolated nodes (allowed, cntm/model.py:198)
    x0 = np.random.default_rng(10).integers(0, 2, 80)
    cntm_case("cntm_er80_isolates", g, 1, 0.3, 0.5, 0.5, x0, 20.0, 21, "PCG64", 17)

    # ---- sample_many_runs: stream threading (serial) and spawn (n_jobs=2) ----------------------
    g = connected(nx.barabasi_albert_graph(100, 3, seed=1))
    p = CNVMParameters(num_opinions=3, network=g, r=R3, r_tilde=RT3)
    states = np.random.default_rng(11).integers(0, 3, (2, 100))
    cv = OpinionShares(3, normalize=True)
    row_ptr, col = csr(p.network)
    common = dict(num_opinions=3, num_agents=100, r=R3, r_tilde=RT3, row_ptr=row_ptr, col=col, x_init=states, t_max=5.0)
    t, x = sample_many_runs(p, states, 5.0, 11, 3, n_jobs=None, rng=Generator(PCG64(21)))
    save("smr_serial_full", t_out=t, x_out=x, seed=21, n_jobs=0, num_runs=3, num_t=11, **common)
    t, x = sample_many_runs(p, states, 5.0, 11, 5, n_jobs=2, collective_variable=cv, rng=Generator(PCG64(22)))
    save("smr_jobs2_runs_cv", t_out=t, x_out=x, seed=22, n_jobs=2, num_runs=5, num_t=11, **common)
    states3 = np.random.default_rng(12).integers(0, 3, (3, 100))
    common["x_init"] = states3
    t, x = sample_many_runs(p, states3, 5.0, 11, 2, n_jobs=2, rng=Generator(PCG64(23)))
    save("smr_jobs2_states_full", t_out=t, x_out=x, seed=23, n_jobs=2, num_runs=2, num_t=11, **common)

    # ---- collective variables (known answers of the reference's own classes) -------------------
    g = connected(nx.barabasi_albert_graph(100, 3, seed=1))
    xs = np.random.default_rng(13).integers(0, 3, (7, 100))
    w = np.random.default_rng(14).random(100) - 0.3
    deg = np.array([d for _, d in g.degree()])
    row_ptr, col = csr(calculate_neighbor_list(g))
    save(
        "cv_known_answers", x=xs, weights=w, degrees=deg, row_ptr=row_ptr, col=col,
        shares=OpinionShares(3)(xs), shares_norm=OpinionShares(3, normalize=True)(xs),
        shares_w=OpinionShares(3, weights=w)(xs), shares_w_norm=OpinionShares(3, True, w, [2, 0])(xs),
        shares_deg=DegreeWeightedOpinionShares(3, g)(xs), shares_deg_norm=DegreeWeightedOpinionShares(3, g, True, 1)(xs),
        by_degree=OpinionSharesByDegree(3, g)(xs), by_degree_norm=OpinionSharesByDegree(3, g, True, [0, 2])(xs),
    )


if __name__ == "__main__":
    main()
