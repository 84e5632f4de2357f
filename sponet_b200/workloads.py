"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d), built identically on every box.

Used by bench.py, the smoke test and the parity tests at full size.  Graph construction follows the
reference's own generator choice (``nx.fast_gnp_random_graph`` for p <= 0.2,
sponet/network_generator.py:39) with isolates patched by one edge, because CNVM forbids isolated
vertices (sponet/cnvm/model.py:36-37).
"""
from __future__ import annotations

import numpy as np

# the reference's standard 3-opinion rates (tests/tests_cnvm/test_cnvm_simulate.py:22-29)
R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)


def er_network(num_agents: int, mean_degree: float = 10.0, seed: int = 12345):
    """G(N, p = mean_degree / (N - 1)); isolated vertex i gets the edge (i, i+1 mod N)."""
    import networkx as nx

    g = nx.fast_gnp_random_graph(num_agents, mean_degree / (num_agents - 1), seed=seed)
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % num_agents)
    return g


def headline_cnvm(num_agents: int = 100_000, seed: int = 12345, alpha: float = 1.0):
    """H: CNVM, M = 3, ER <d> = 10, fixture rates (r_imit = 2, r_noise = 0.6), alpha = 1.

    Returns (params, x_init, events_per_unit_time).  At N = 1e5: nnz = 998 782, lambda = 2.6e5.
    (alpha != 1 gives the degree-weighted agent selection: alias tables on the hot path.)
    """
    from .cnvm import CNVMParameters
    from .utils import calculate_neighbor_list

    g = er_network(num_agents, 10.0, seed)
    nl = calculate_neighbor_list(g)
    params = CNVMParameters(num_opinions=3, network=nl, r=R3, r_tilde=RT3, alpha=alpha)
    x_init = np.random.default_rng(0).integers(0, 3, num_agents).astype(np.uint8)
    # cnvm/model.py:257: r_imit * sum_i d_i^(1 - alpha) + r_noise * N
    rate = params.r_imit * float(np.sum(np.array([len(nb) for nb in nl], dtype=float) ** (1.0 - alpha))) + params.r_noise * num_agents
    return params, x_init, rate


def complete_cnvm(num_agents: int = 1000):
    """Config 1: CNVM, M = 2, complete graph, r = 1, r_tilde = 0.01 (benchmarks/benchmark_cnvm.py:13-15)."""
    from .cnvm import CNVMParameters

    params = CNVMParameters(num_opinions=2, num_agents=num_agents, r=1.0, r_tilde=0.01)
    x_init = np.random.default_rng(0).integers(0, 2, num_agents).astype(np.uint8)
    return params, x_init, num_agents * (params.r_imit + params.r_noise)


def ba_cntm(num_agents: int = 100_000, m: int = 5, seed: int = 1):
    """Config 3: CNTM on Barabasi-Albert, r = 1, r_tilde = 0.2, thresholds 0.5 / 0.5."""
    import networkx as nx

    from .cntm import CNTMParameters

    g = nx.barabasi_albert_graph(num_agents, m, seed=seed)
    params = CNTMParameters(network=g, r=1.0, r_tilde=0.2, threshold_01=0.5, threshold_10=0.5)
    x_init = np.random.default_rng(0).integers(0, 2, num_agents).astype(np.uint8)
    return params, x_init, num_agents * (params.r + params.r_tilde)


def sis_ws(num_agents: int = 1_000_000, infection_rate: float = 1.0, seed: int = 7):
    """Config 4: SIS written as a CNVM (examples/SIS_model.py:77-80,108-112) on a Watts-Strogatz ring
    (k = 10, p = 0.1): r = [[0, lambda], [0, 0]], r_tilde = [[0, 0], [1, 0]], 30 % infected initially."""
    import networkx as nx

    from .cnvm import CNVMParameters
    from .utils import calculate_neighbor_list

    g = nx.connected_watts_strogatz_graph(num_agents, 10, 0.1, seed=seed) if num_agents <= 20_000 else \
        nx.watts_strogatz_graph(num_agents, 10, 0.1, seed=seed)  # k = 10 ring + rewiring never isolates a node
    r = np.array([[0.0, infection_rate], [0.0, 0.0]])
    rt = np.array([[0.0, 0.0], [1.0, 0.0]])
    params = CNVMParameters(num_opinions=2, network=calculate_neighbor_list(g), r=r, r_tilde=rt, alpha=1.0)
    x_init = np.zeros(num_agents, dtype=np.uint8)
    x_init[: int(0.3 * num_agents)] = 1
    np.random.default_rng(0).shuffle(x_init)
    rate = params.r_imit * num_agents + params.r_noise * num_agents
    return params, x_init, rate
