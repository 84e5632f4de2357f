"""Reference CV samples for the statistical parity tests of native mode (build container only).

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_stats_golden.py [name ...]

For each configuration the UNMODIFIED reference runs >= 1e4 realizations through
`sample_many_runs(..., n_jobs=8)` with OpinionShares as collective variable; the per-run counts
[R, T, M] are stored in tests/golden/stats_*.npz together with the inputs (uint16, or uint32 when
N > 65535).  Seeds are fixed integers (SEEDS), so a fixture is reproducible bit for bit with the same
numpy / numba / networkx versions and n_jobs.

Small shapes (N = 300-500) carry their CSR; the BASELINE-size shapes (config 2: ER N = 1e4; H: ER N = 1e5;
config 3: CNTM on BA N = 1e4 / 1e5) rebuild their network from `sponet_b200.workloads` (same generator,
same seed, on every box) and store a CRC of the CSR so that a drifting generator is noticed.
"""
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")
sys.path.insert(1, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import zlib  # noqa: E402

import networkx as nx  # noqa: E402
import numpy as np  # noqa: E402

from sponet import CNTMParameters, CNVMParameters, OpinionShares, sample_many_runs  # noqa: E402
from sponet.utils import calculate_neighbor_list  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)
RUNS = 12_000
SEEDS = {"cnvm_er500": 101, "cnvm_ba400_alpha0": 102, "cnvm_complete300": 103, "cntm_er400": 104,
         "cnvm_er1e4": 201, "cnvm_er1e5": 202, "cntm_ba1e4": 203, "cntm_ba1e5": 204,
         "cnvm_pernet_er1000": 301, "cnvm_pernet_ba1000": 302, "cnvm_pernet_ws1000": 303, "cnvm_pernet_reg1000": 304, "cnvm_pernet_ba1000_alpha0": 305}
ONLY = set(sys.argv[1:])


def csr(nl):
    deg = np.array([len(nb) for nb in nl])
    rp = np.zeros(len(nl) + 1, dtype=np.int32)
    np.cumsum(deg, out=rp[1:])
    return rp, np.concatenate([np.asarray(nb, dtype=np.int32) for nb in nl])


def patch(g):
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % g.number_of_nodes())
    return g


def run(name, params, M, x0, t_max, T, extra, runs=RUNS):
    if ONLY and name not in ONLY:
        return
    cv = OpinionShares(M)
    _, x = sample_many_runs(params, x0, t_max, T, runs, n_jobs=8, collective_variable=cv, rng=np.random.default_rng(SEEDS[name]))
    dt = np.uint16 if x0.size < 65536 else np.uint32
    assert np.array_equal(x, np.round(x)) and x.max() <= np.iinfo(dt).max
    np.savez_compressed(os.path.join(OUT, f"stats_{name}.npz"), counts=x.astype(dt), x_init=x0.astype(np.uint8),
                        t_max=t_max, num_t=T, num_opinions=M, seed=SEEDS[name], **extra)
    print(name, x.shape, "mean at t_max:", x[:, -1].mean(0), flush=True)


def crc(rp, col):
    return zlib.crc32(np.ascontiguousarray(col).tobytes(), zlib.crc32(np.ascontiguousarray(rp).tobytes()))


# CNVM on a network, alpha = 1 (uniform agent choice)
g = patch(nx.fast_gnp_random_graph(500, 10 / 499, seed=1))
nl = calculate_neighbor_list(g)
rp, col = csr(nl)
x0 = np.random.default_rng(0).integers(0, 3, 500)
run("cnvm_er500", CNVMParameters(num_opinions=3, network=nl, r=R3, r_tilde=RT3), 3, x0, 5.0, 6,
    dict(model="cnvm", row_ptr=rp, col=col, r=R3, r_tilde=RT3, alpha=1.0))

# CNVM on a heterogeneous network with alpha = 0 (alias tables)
g = nx.barabasi_albert_graph(400, 3, seed=2)
nl = calculate_neighbor_list(g)
rp, col = csr(nl)
x0 = np.random.default_rng(1).integers(0, 3, 400)
run("cnvm_ba400_alpha0", CNVMParameters(num_opinions=3, network=nl, r=R3, r_tilde=RT3, alpha=0.0), 3, x0, 2.0, 5,
    dict(model="cnvm", row_ptr=rp, col=col, r=R3, r_tilde=RT3, alpha=0.0))

# CNVM complete graph (BASELINE config 1 rates, smaller N)
x0 = np.random.default_rng(2).integers(0, 2, 300)
r2, rt2 = np.array([[0, 1.0], [1.0, 0]]), np.array([[0, 0.01], [0.01, 0]])
run("cnvm_complete300", CNVMParameters(num_opinions=2, num_agents=300, r=1.0, r_tilde=0.01), 2, x0, 20.0, 5,
    dict(model="cnvm_complete", r=r2, r_tilde=rt2, alpha=1.0))

# CNTM
g = nx.fast_gnp_random_graph(400, 8 / 399, seed=3)
nl = calculate_neighbor_list(g)
rp, col = csr(nl)
x0 = np.random.default_rng(3).integers(0, 2, 400)
run("cntm_er400", CNTMParameters(network=g, r=1.0, r_tilde=0.2, threshold_01=0.5, threshold_10=0.5), 2, x0, 10.0, 6,
    dict(model="cntm", row_ptr=rp, col=col, r=1.0, r_tilde=0.2, threshold_01=0.5, threshold_10=0.5))

# ---- BASELINE-size shapes: networks come from sponet_b200.workloads (rebuilt by the test, CRC-checked) ----
from sponet_b200 import workloads  # noqa: E402  (host-side generators only; no GPU code is touched)

for name, n, t_max, T in (("cnvm_er1e4", 10_000, 10.0, 6), ("cnvm_er1e5", 100_000, 2.0, 5)):
    if ONLY and name not in ONLY:
        continue
    g = workloads.er_network(n, 10.0, 12345)
    nl = calculate_neighbor_list(g)
    rp, col = csr(nl)
    x0 = np.random.default_rng(0).integers(0, 3, n)
    run(name, CNVMParameters(num_opinions=3, network=nl, r=R3, r_tilde=RT3), 3, x0, t_max, T,
        dict(model="cnvm", workload="headline_cnvm", num_agents=n, csr_crc=crc(rp, col), r=R3, r_tilde=RT3, alpha=1.0), runs=10_000)

for name, n, t_max, T in (("cntm_ba1e4", 10_000, 10.0, 6), ("cntm_ba1e5", 100_000, 2.0, 5)):
    if ONLY and name not in ONLY:
        continue
    g = nx.barabasi_albert_graph(n, 5, seed=1)
    nl = calculate_neighbor_list(g)
    rp, col = csr(nl)
    x0 = np.random.default_rng(0).integers(0, 2, n)
    run(name, CNTMParameters(network=g, r=1.0, r_tilde=0.2, threshold_01=0.5, threshold_10=0.5), 2, x0, t_max, T,
        dict(model="cntm", workload="ba_cntm", num_agents=n, csr_crc=crc(rp, col), r=1.0, r_tilde=0.2, threshold_01=0.5,
             threshold_10=0.5), runs=10_000)


# ---- a NEW network for every realization (params.network_generator; sponet/cnvm/model.py:95-96) ----
# The reference's own generators (networkx); the device-side generators of this repo must give the same distribution of
# the collective variable -- the graph-to-graph variation is part of it at N = 1000.
import sponet.network_generator as ng  # noqa: E402

x0 = np.random.default_rng(4).integers(0, 3, 1000)
run("cnvm_pernet_er1000",
    CNVMParameters(num_opinions=3, network_generator=ng.ErdosRenyiGenerator(1000, 0.02, rng=np.random.default_rng(11)), r=R3, r_tilde=RT3),
    3, x0, 4.0, 5, dict(model="cnvm_pernet", generator="er", num_agents=1000, p=0.02, r=R3, r_tilde=RT3, alpha=1.0))
run("cnvm_pernet_ba1000",
    CNVMParameters(num_opinions=3, network_generator=ng.BarabasiAlbertGenerator(1000, 3, rng=np.random.default_rng(12)), r=R3, r_tilde=RT3),
    3, x0, 4.0, 5, dict(model="cnvm_pernet", generator="ba", num_agents=1000, m=3, r=R3, r_tilde=RT3, alpha=1.0))
run("cnvm_pernet_ws1000",
    CNVMParameters(num_opinions=3, network_generator=ng.WattsStrogatzGenerator(1000, 6, 0.2, rng=np.random.default_rng(13)), r=R3, r_tilde=RT3),
    3, x0, 4.0, 5, dict(model="cnvm_pernet", generator="ws", num_agents=1000, k=6, p=0.2, r=R3, r_tilde=RT3, alpha=1.0))
run("cnvm_pernet_reg1000",
    CNVMParameters(num_opinions=3, network_generator=ng.RandomRegularGenerator(1000, 6, rng=np.random.default_rng(14)), r=R3, r_tilde=RT3),
    3, x0, 4.0, 5, dict(model="cnvm_pernet", generator="regular", num_agents=1000, d=6, r=R3, r_tilde=RT3, alpha=1.0))
run("cnvm_pernet_ba1000_alpha0",
    CNVMParameters(num_opinions=3, network_generator=ng.BarabasiAlbertGenerator(1000, 3, rng=np.random.default_rng(15)), r=R3, r_tilde=RT3,
                   alpha=0.0),
    3, x0, 0.5, 5, dict(model="cnvm_pernet", generator="ba", num_agents=1000, m=3, r=R3, r_tilde=RT3, alpha=0.0))
