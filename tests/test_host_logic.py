"""CPU: host-side logic of sponet_b200 against known answers produced by the unmodified reference
(tests/golden/host_logic.json, made by tests/golden/make_host_golden.py) and against the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

with open(os.path.join(GOLDEN, "host_logic.json")) as fh:
    KA = json.load(fh)


@pytest.mark.parametrize("case", KA["rates_style1"])
def test_rates_style1(case):
    from sponet_b200 import CNVMParameters
    from sponet_b200.cnvm.parameters import convert_rate_from_cnvm

    p = CNVMParameters(case["num_opinions"], r=case["r_in"], r_tilde=case["r_tilde_in"], num_agents=10)
    assert p.num_opinions == case["M"]
    np.testing.assert_array_equal(p.r, case["r"])
    np.testing.assert_array_equal(p.r_tilde, case["r_tilde"])
    assert p.r_imit == case["r_imit"] and p.r_noise == case["r_noise"]
    np.testing.assert_array_equal(p.prob_imit, case["prob_imit"])
    np.testing.assert_array_equal(p.prob_noise, case["prob_noise"])
    r, rt = convert_rate_from_cnvm(p)  # round trip
    np.testing.assert_allclose(r, case["r"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(rt, case["r_tilde"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", KA["rates_style2"])
def test_rates_style2(case):
    from sponet_b200 import CNVMParameters

    p = CNVMParameters(case["num_opinions"], num_agents=10, r_imit=case["r_imit"], r_noise=case["r_noise"],
                       prob_imit=case["prob_imit_in"], prob_noise=case["prob_noise_in"])
    assert p.num_opinions == case["M"]
    for k in ("r", "r_tilde", "prob_imit", "prob_noise"):
        np.testing.assert_array_equal(getattr(p, k), case[k])


@pytest.mark.parametrize("case", KA["rate_errors"])
def test_parameter_errors(case):
    from sponet_b200 import CNVMParameters

    assert case["raises"] == "ValueError"
    with pytest.raises(ValueError):
        CNVMParameters(**case["kwargs"])


def test_change_rates_and_network_precedence():
    import networkx as nx

    from sponet_b200 import CNVMParameters

    g = nx.star_graph(4)
    p = CNVMParameters(num_opinions=2, num_agents=99, network=g, r=1, r_tilde=0.1)
    assert p.num_agents == 5 and len(p.network) == 5 and p.network[0].dtype == np.int32  # network beats num_agents
    p.change_rates(r=[[0, 2], [4, 0]])
    assert p.r_imit == 4 and p.prob_imit[0, 1] == 0.5 and p.r_tilde[0, 1] == 0.1
    p.set_network(None)
    assert p.network is None and p.num_agents == 5
    adj = [[1], [0, 2], [1]]
    p2 = CNVMParameters(num_opinions=2, network=adj, r=1, r_tilde=0)
    assert p2.num_agents == 3 and p2.r_noise == 0 and not p2.prob_noise.any()
    assert sorted(p2.get_network().edges) == [(0, 1), (1, 2)]


@pytest.mark.parametrize("case", KA["argmatch"])
def test_argmatch(case, oracle):
    from sponet_b200.utils import argmatch

    got = argmatch(np.array(case["x_ref"]), np.array(case["x"]))
    np.testing.assert_array_equal(got, case["out"])
    np.testing.assert_array_equal(oracle.argmatch(case["x_ref"], case["x"]), case["out"])


@pytest.mark.parametrize("case", KA["counts_from_shares"])
def test_counts_from_shares(case):
    from sponet_b200.utils import counts_from_shares

    np.testing.assert_array_equal(counts_from_shares(case["shares"], case["num_agents"]), case["counts"])


@pytest.mark.parametrize("case", KA["split_runs"])
def test_split_runs(case):
    from sponet_b200.multiprocessing import _split_runs

    np.testing.assert_array_equal(_split_runs(case["num_runs"], case["num_chunks"]), case["chunks"])


@pytest.mark.parametrize("case", KA["mask_dups"])
def test_mask_subsequent_duplicates(case):
    from sponet_b200.utils import mask_subsequent_duplicates

    np.testing.assert_array_equal(mask_subsequent_duplicates(np.array(case["x"])), case["mask"])
    with pytest.raises(ValueError):
        mask_subsequent_duplicates(np.zeros((2, 2, 2)))


@pytest.mark.parametrize("case", KA["t_eval"])
def test_t_eval_grid(case):
    from sponet_b200.utils import t_eval_to_ndarray

    np.testing.assert_array_equal(t_eval_to_ndarray(case["t_eval"], case["t_max"]), case["out"])


@pytest.mark.parametrize("case", KA["t_eval_errors"])
def test_t_eval_errors(case):
    from sponet_b200.utils import t_eval_to_ndarray

    kind, msg = case["raises"]
    assert kind == "ValueError"
    with pytest.raises(ValueError) as err:
        t_eval_to_ndarray(case["t_eval"], case["t_max"])
    assert str(err.value) == msg


@pytest.mark.parametrize("case", KA["alias"])
def test_alias_table_library_and_oracle(case, oracle):
    """sponet_build_alias_table runs on the host (no GPU needed): identical tables to the reference."""
    from sponet_b200.sampling import build_alias_table

    w = np.array([float.fromhex(h) for h in case["weights"]])
    prob_ref = np.array([float.fromhex(h) for h in case["prob"]])
    for impl in (build_alias_table, oracle.build_alias_table):
        prob, alias = impl(w)
        np.testing.assert_array_equal(prob, prob_ref)
        np.testing.assert_array_equal(alias, case["alias"])


def test_cnvm_event_rates_sequential_sum(oracle):
    """next_event_rate uses numba's left-to-right np.sum, not numpy's pairwise sum."""
    from sponet_b200 import CNVMParameters, engine

    rng = np.random.default_rng(0)
    n = 5000
    adj = [np.array([(i + 1) % n], dtype=np.int32) for i in range(n)]
    p = CNVMParameters(num_opinions=2, network=adj, r=1.3, r_tilde=0.07, alpha=0.37)
    da = rng.integers(1, 40, n).astype(float) ** (1 - 0.37)
    rate, noise = engine.cnvm_event_rates(p, da, complete=False)
    total = 0.0
    for v in da:
        total += v
    assert rate == 1.0 / (p.r_imit * total + p.r_noise * n)
    assert noise == p.r_noise * n * rate
    rate_c, noise_c = engine.cnvm_event_rates(p, np.array([7.0]), complete=True)
    assert rate_c == 1.0 / (p.r_imit * 7.0 + p.r_noise) / n
    assert noise_c == p.r_noise / (p.r_noise + p.r_imit * 7.0)
    p0 = CNVMParameters(num_opinions=2, network=adj, r=0, r_tilde=0)
    with pytest.raises(ZeroDivisionError):
        engine.cnvm_event_rates(p0, da, complete=False)


def test_neighbor_list_and_csr():
    import networkx as nx

    from sponet_b200.utils import calculate_neighbor_list, neighbor_list_to_csr

    nl = calculate_neighbor_list(nx.star_graph(3))
    assert [sorted(x.tolist()) for x in nl] == [[1, 2, 3], [0], [0], [0]] and nl[0].dtype == np.int32
    row_ptr, col = neighbor_list_to_csr(nl)
    np.testing.assert_array_equal(row_ptr, [0, 3, 4, 5, 6])
    np.testing.assert_array_equal(col, np.concatenate(nl))
    row_ptr, col = neighbor_list_to_csr([np.zeros(0, dtype=np.int32)] * 3)
    assert row_ptr.tolist() == [0, 0, 0, 0] and col.size == 0


def test_isolated_vertices_rejected_for_cnvm():
    from sponet_b200 import CNVM, CNVMParameters

    adj = [np.array([1], dtype=np.int32), np.array([0], dtype=np.int32), np.zeros(0, dtype=np.int32)]
    with pytest.raises(ValueError, match="Isolated vertices"):
        CNVM(CNVMParameters(num_opinions=2, network=adj, r=1, r_tilde=0.1))


def test_degree_alpha_matches_reference_formula():
    import networkx as nx

    from sponet_b200 import CNVM, CNVMParameters

    g = nx.barabasi_albert_graph(50, 2, seed=0)
    m = CNVM(CNVMParameters(num_opinions=2, network=g, r=1, r_tilde=0.1, alpha=0.25))
    deg = np.array([d for _, d in g.degree()])
    np.testing.assert_array_equal(m.degree_alpha, deg ** (1 - 0.25))
    mc = CNVM(CNVMParameters(num_opinions=2, num_agents=30, r=1, r_tilde=0.1, alpha=0.5))
    np.testing.assert_array_equal(mc.degree_alpha, np.ones(30) * 29 ** 0.5)


def test_log1p_restatement_matches_libm():
    """csrc/sponet_log1p.h reproduces the host libm bit for bit (needed for bit-exact ziggurat tails).
    Compiled here with gcc -ffp-contract=off; skipped if the host libm is not the glibc FMA build."""
    import ctypes
    import subprocess
    import tempfile

    root = os.path.dirname(GOLDEN).rsplit(os.sep, 1)[0]
    src = '#include "%s/sponet_b200/csrc/sponet_log1p.h"\ndouble f(double x){return sponet_log1p(x);}\n' % root
    with tempfile.TemporaryDirectory() as d:
        c, so = os.path.join(d, "l.c"), os.path.join(d, "l.so")
        open(c, "w").write(src)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", c, "-o", so, "-lm"], check=True)
        lib = ctypes.CDLL(so)
        lib.f.restype, lib.f.argtypes = ctypes.c_double, [ctypes.c_double]
        libm = ctypes.CDLL("libm.so.6")
        libm.log1p.restype, libm.log1p.argtypes = ctypes.c_double, [ctypes.c_double]
        u = np.random.default_rng(1).integers(0, 2**53, 200_000).astype(np.float64) * 2.0**-53
        u[:1000] *= 1e-9
        bad = sum(lib.f(-x) != libm.log1p(-x) for x in u)
        if bad > 100:
            pytest.skip(f"host libm log1p is not the glibc FMA variant ({bad} mismatches)")
        assert bad == 0
