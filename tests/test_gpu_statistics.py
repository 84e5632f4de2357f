"""GPU: native-RNG mode against the reference's numba path, statistically.

Fixtures tests/golden/stats_*.npz hold OpinionShares counts of 12 000 runs of the UNMODIFIED reference
(tests/golden/make_stats_golden.py).  Native mode samples 24 000 runs of the same configuration.  For
every output time k >= 1 and every opinion m:

  * z-test on the mean:      |mean_a - mean_b| <= Z * sqrt(var_a/n_a + var_b/n_b)
  * z-test on the variance:  |var_a - var_b|  <= Z * sqrt(se_a^2 + se_b^2), se^2 = (m4 - var^2)/n
  * two-sample Kolmogorov-Smirnov: p-value >= ALPHA / (number of comparisons)   (Bonferroni)

with Z = 4.75 (two-sided 2e-6 per comparison; <= 2e-4 family-wise over <= 100 comparisons per test)
and ALPHA = 1e-3 family-wise for KS.  A kernel with a wrong rate, branch probability, neighbour
choice or acceptance table fails these by tens of sigma (checked by perturbing r_tilde by 5 %).
"""
import numpy as np
import pytest
from scipy import stats

from conftest import csr_to_neighbor_list, load_golden

pytestmark = pytest.mark.gpu

Z = 4.75
ALPHA = 1e-3
NATIVE_RUNS = 24_000


def _crc(rp, col):
    import zlib

    return zlib.crc32(np.ascontiguousarray(col).tobytes(), zlib.crc32(np.ascontiguousarray(rp).tobytes()))


def _workload_params(g, scale_noise):
    """BASELINE-size fixtures do not carry their CSR: the network is rebuilt from sponet_b200.workloads (same
    generator and seed as tests/golden/make_stats_golden.py) and its CRC compared with the stored one."""
    from sponet_b200 import CNTMParameters, CNVMParameters, workloads
    from sponet_b200.utils import calculate_neighbor_list, neighbor_list_to_csr

    n = int(g["num_agents"])
    if str(g["workload"]) == "headline_cnvm":
        nl = calculate_neighbor_list(workloads.er_network(n, 10.0, 12345))
        params = CNVMParameters(num_opinions=3, network=nl, r=g["r"], r_tilde=g["r_tilde"] * scale_noise)
    else:
        import networkx as nx

        net = nx.barabasi_albert_graph(n, 5, seed=1)
        nl = calculate_neighbor_list(net)
        params = CNTMParameters(net, float(g["r"]), float(g["r_tilde"]) * scale_noise, float(g["threshold_01"]),
                                float(g["threshold_10"]))
    rp, col = neighbor_list_to_csr(nl)
    assert _crc(rp.astype(np.int32), col.astype(np.int32)) == int(g["csr_crc"]), "network generator drifted from the fixture's"
    return params


def _native_counts(g, runs, seed, scale_noise=1.0):
    import networkx as nx

    from sponet_b200 import CNTMParameters, CNVMParameters, OpinionShares, sample_many_runs

    model, M = str(g["model"]), int(g["num_opinions"])
    if model == "cnvm_pernet":  # a new network per realization, generated on the device
        from sponet_b200 import BarabasiAlbertGenerator, ErdosRenyiGenerator

        n = int(g["num_agents"])
        from sponet_b200 import RandomRegularGenerator, WattsStrogatzGenerator

        kind, grng = str(g["generator"]), np.random.default_rng(seed + 1)
        gen = (ErdosRenyiGenerator(n, float(g["p"]), rng=grng) if kind == "er"
               else BarabasiAlbertGenerator(n, int(g["m"]), rng=grng) if kind == "ba"
               else WattsStrogatzGenerator(n, int(g["k"]), float(g["p"]), rng=grng) if kind == "ws"
               else RandomRegularGenerator(n, int(g["d"]), rng=grng))
        params = CNVMParameters(num_opinions=M, network_generator=gen, r=g["r"], r_tilde=g["r_tilde"] * scale_noise,
                                alpha=float(g["alpha"]))
    elif "workload" in g:
        params = _workload_params(g, scale_noise)
    elif model == "cntm":
        nl = csr_to_neighbor_list(g["row_ptr"], g["col"])
        net = nx.Graph()
        net.add_nodes_from(range(len(nl)))
        net.add_edges_from((i, int(j)) for i, nb in enumerate(nl) for j in nb if i < j)
        params = CNTMParameters(net, float(g["r"]), float(g["r_tilde"]) * scale_noise, float(g["threshold_01"]),
                                float(g["threshold_10"]))
    elif model == "cnvm_complete":
        params = CNVMParameters(num_opinions=M, num_agents=g["x_init"].size, r=g["r"], r_tilde=g["r_tilde"] * scale_noise)
    else:
        params = CNVMParameters(num_opinions=M, network=csr_to_neighbor_list(g["row_ptr"], g["col"]), r=g["r"],
                                r_tilde=g["r_tilde"] * scale_noise, alpha=float(g["alpha"]))
    _, x = sample_many_runs(params, g["x_init"], float(g["t_max"]), int(g["num_t"]), runs,
                            collective_variable=OpinionShares(M), rng=np.random.default_rng(seed))
    return x


def _compare(ref, nat):
    """Returns the worst z (mean), worst z (variance), smallest KS p-value and the comparison count."""
    T, M = ref.shape[1:]
    zm = zv = 0.0
    pmin, n_cmp = 1.0, 0
    for k in range(1, T):
        for m in range(M):
            a, b = ref[:, k, m].astype(float), nat[:, k, m].astype(float)
            va, vb = a.var(), b.var()
            if va == 0 and vb == 0:
                assert a.mean() == b.mean()
                continue
            n_cmp += 1
            zm = max(zm, abs(a.mean() - b.mean()) / np.sqrt(va / a.size + vb / b.size))
            m4a, m4b = ((a - a.mean()) ** 4).mean(), ((b - b.mean()) ** 4).mean()
            se = np.sqrt((m4a - va**2) / a.size + (m4b - vb**2) / b.size)
            zv = max(zv, abs(va - vb) / se)
            pmin = min(pmin, stats.ks_2samp(a, b).pvalue)
    return zm, zv, pmin, n_cmp


@pytest.mark.parametrize("name", ["cnvm_er500", "cnvm_ba400_alpha0", "cnvm_complete300", "cntm_er400"])
def test_native_mode_matches_reference_distribution(name):
    g = load_golden("stats_" + name)
    ref = g["counts"]
    nat = _native_counts(g, NATIVE_RUNS, seed=2026)
    assert nat.shape == (NATIVE_RUNS,) + ref.shape[1:]
    np.testing.assert_array_equal(nat[:, 0], np.broadcast_to(ref[0, 0], nat[:, 0].shape))  # t = 0: the initial state
    assert (nat.sum(-1) == g["x_init"].size).all()
    zm, zv, pmin, n_cmp = _compare(ref, nat)
    print(f"{name}: worst z(mean)={zm:.2f} z(var)={zv:.2f} min KS p={pmin:.3g} over {n_cmp} comparisons")
    assert zm <= Z, f"means differ: z = {zm:.2f}"
    assert zv <= Z, f"variances differ: z = {zv:.2f}"
    assert pmin >= ALPHA / n_cmp, f"KS rejects: p = {pmin:.3g}"


@pytest.mark.parametrize("name", ["cnvm_pernet_er1000", "cnvm_pernet_ba1000", "cnvm_pernet_ws1000", "cnvm_pernet_reg1000",
                                  "cnvm_pernet_ba1000_alpha0"])
def test_per_run_networks_match_the_reference_distribution(name):
    """`params.network_generator`: the reference draws a networkx graph per realization (its ErdosRenyiGenerator /
    BarabasiAlbertGenerator / WattsStrogatzGenerator / RandomRegularGenerator, 1.2e4 realizations through its own `sample_many_runs`); here the graphs of all realizations are
    generated on the device.  At N = 1000 the variation from graph to graph is a visible part of the CV's distribution, so
    means, variances and KS test the generated ensemble of graphs as well as the chains on them."""
    g = load_golden("stats_" + name)
    ref = g["counts"].astype(np.int64)
    nat = _native_counts(g, NATIVE_RUNS, seed=777)
    assert nat.shape == (NATIVE_RUNS,) + ref.shape[1:]
    np.testing.assert_array_equal(nat[:, 0], np.broadcast_to(ref[0, 0], nat[:, 0].shape))
    zm, zv, pmin, n_cmp = _compare(ref, nat)
    print(f"{name}: worst z(mean)={zm:.2f} z(var)={zv:.2f} min KS p={pmin:.3g} over {n_cmp} comparisons")
    assert zm <= Z, f"means differ: z = {zm:.2f}"
    assert zv <= Z, f"variances differ: z = {zv:.2f}"
    assert pmin >= ALPHA / n_cmp, f"KS rejects: p = {pmin:.3g}"


@pytest.mark.parametrize("name", ["cnvm_er1e4", "cnvm_er1e5", "cntm_ba1e4", "cntm_ba1e5"])
def test_native_mode_matches_reference_at_baseline_sizes(name):
    """The same comparison on the BASELINE.json shapes: config 2 (CNVM M=3, ER N=1e4), H (ER N=1e5), config 3
    (CNTM on BA m=5, N=1e4 and N=1e5), 1e4 reference runs each (unmodified numba path, n_jobs=8) against 2e4
    native runs.  These sizes exercise what the small fixtures cannot: the neighbour sampler table at lg=5,
    agent selection quantised to N / 2^32, the Bloom-filter sizing, the two-warps-per-chain schedule, hub rows
    of ~10^3 neighbours (BA), and the native snapshot rule (state AT t_eval[k]) against the reference's
    nearest-event rule."""
    g = load_golden("stats_" + name)
    ref = g["counts"].astype(np.int64)
    runs = 20_000
    nat = _native_counts(g, runs, seed=4242)
    assert nat.shape == (runs,) + ref.shape[1:]
    np.testing.assert_array_equal(nat[:, 0], np.broadcast_to(ref[0, 0], nat[:, 0].shape))
    assert (nat.sum(-1) == g["x_init"].size).all()
    zm, zv, pmin, n_cmp = _compare(ref, nat)
    print(f"{name}: worst z(mean)={zm:.2f} z(var)={zv:.2f} min KS p={pmin:.3g} over {n_cmp} comparisons")
    assert zm <= Z, f"means differ: z = {zm:.2f}"
    assert zv <= Z, f"variances differ: z = {zv:.2f}"
    assert pmin >= ALPHA / n_cmp, f"KS rejects: p = {pmin:.3g}"


def test_statistical_test_has_power_at_baseline_size():
    """At N = 1e4 the comparison must still reject a 2 % perturbation of the noise rates."""
    g = load_golden("stats_cnvm_er1e4")
    nat = _native_counts(g, 20_000, seed=8, scale_noise=1.02)
    zm, _, _, _ = _compare(g["counts"].astype(np.int64), nat)
    assert zm > 2 * Z, f"test is too weak: z = {zm:.2f}"


def test_statistical_test_has_power():
    """The same comparison must REJECT a 5 % perturbation of the noise rates."""
    g = load_golden("stats_cnvm_er500")
    nat = _native_counts(g, NATIVE_RUNS, seed=7, scale_noise=1.05)
    zm, _, _, _ = _compare(g["counts"], nat)
    assert zm > 2 * Z, f"test is too weak: z = {zm:.2f}"
