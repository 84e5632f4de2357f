"""GPU: initial-state samplers (sponet_sample_states) and the well-mixed share sampler on the count-only kernel
(reference sponet/states.py:45-164, sponet/cnvm/approximations/stochastic_approximation.py).

Bars: the kernel equals its sequential statement (oracle_sample_states) bit for bit; the defining properties of the
reference samplers hold exactly (opinion counts of every state == counts_from_shares(target)); arrangement and
opinions are uniform (z-tests, tolerances below); the share sampler agrees with the agent-level complete-graph
kernel it approximates."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,M,S", [(1000, 3, 5), (257, 2, 64), (5000, 7, 33), (64, 256, 3)])
def test_sampler_kernel_matches_oracle(oracle, n, M, S):
    import ctypes as C

    from sponet_b200 import _lib

    rng = np.random.default_rng(n)
    seed = 99
    out = np.empty((S, n), dtype=np.uint8)
    _lib.check(_lib.lib().sponet_sample_states(n, M, None, 0, S, seed, _lib.ptr(out), None))
    np.testing.assert_array_equal(out, oracle.sample_states(n, M, None, S, seed))
    counts = rng.multinomial(n, np.ones(M) / M, size=S).astype(np.int64)
    _lib.check(_lib.lib().sponet_sample_states(n, M, _lib.ptr(counts), 0, S, seed, _lib.ptr(out), None))
    np.testing.assert_array_equal(out, oracle.sample_states(n, M, counts, S, seed))
    for s in range(S):
        np.testing.assert_array_equal(np.bincount(out[s], minlength=M), counts[s])
    one = counts[:1]
    _lib.check(_lib.lib().sponet_sample_states(n, M, _lib.ptr(one), 1, S, seed, _lib.ptr(out), None))
    np.testing.assert_array_equal(out, oracle.sample_states(n, M, one, S, seed))
    del C


def test_reference_sampler_contracts():
    from sponet_b200 import (sample_states_target_shares, sample_states_uniform, sample_states_uniform_shares)
    from sponet_b200.utils import counts_from_shares

    rng = np.random.default_rng(0)
    n = 2000
    x = sample_states_target_shares(n, [0.2, 0.5, 0.3], 50, rng=rng)
    assert x.shape == (50, n) and x.dtype == np.int64
    target = counts_from_shares([0.2, 0.5, 0.3], n)
    for row in x:
        np.testing.assert_array_equal(np.bincount(row, minlength=3), target)
    assert np.unique(x, axis=0).shape[0] == 50  # unique=True
    # arrangement is uniform: the mean position of the opinion-0 agents is (n - 1) / 2 +- 5 sigma
    pos = np.concatenate([np.flatnonzero(row == 0) for row in x]).astype(float)
    sigma = np.sqrt((n * n - 1) / 12.0 / pos.size)
    assert abs(pos.mean() - (n - 1) / 2) < 5 * sigma
    # a single state is squeezed
    assert sample_states_target_shares(n, [0.5, 0.5], rng=rng).shape == (n,)

    u = sample_states_uniform(n, 4, 100, rng=rng)
    assert u.shape == (100, n) and u.min() == 0 and u.max() == 3
    freq = np.bincount(u.ravel(), minlength=4) / u.size
    assert np.abs(freq - 0.25).max() < 5 * np.sqrt(0.25 * 0.75 / u.size)

    us = sample_states_uniform_shares(n, 3, 400, rng=rng)
    shares = np.stack([np.bincount(r, minlength=3) for r in us]) / n
    # Dirichlet(1,1,1): every share has mean 1/3 and variance 1/18
    assert np.abs(shares.mean(0) - 1 / 3).max() < 5 * np.sqrt(1 / 18 / 400)
    assert np.abs(shares.var(0) - 1 / 18).max() < 0.02


def test_device_states_feed_sample_many_runs():
    import torch

    from sponet_b200 import CNVMParameters, OpinionShares, sample_many_runs, sample_states_target_shares, workloads

    params, _, _ = workloads.headline_cnvm(3000)
    d = sample_states_target_shares(3000, [0.1, 0.3, 0.6], 7, rng=np.random.default_rng(1), device=True)
    assert isinstance(d, torch.Tensor) and d.is_cuda and d.shape == (7, 3000)
    cv = OpinionShares(3)
    t, c = sample_many_runs(params, d, 1.0, 4, 9, collective_variable=cv, rng=np.random.default_rng(2))
    t2, c2 = sample_many_runs(params, d.cpu().numpy(), 1.0, 4, 9, collective_variable=cv, rng=np.random.default_rng(2))
    assert c.shape == (7, 9, 4, 3)
    np.testing.assert_array_equal(c, c2)
    np.testing.assert_array_equal(c[:, :, 0], np.broadcast_to(np.array([300.0, 900.0, 1800.0]), (7, 9, 3)))
    del CNVMParameters


def test_stochastic_approximation_matches_complete_graph_ensemble():
    """The share sampler (count-only kernel) against the agent-level kernel on the complete graph, 2e4 runs each:
    means within 5 standard errors at every output time."""
    from sponet_b200 import CNVMParameters, OpinionShares, sample_many_runs, sample_stochastic_approximation

    R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
    RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)
    N, R, T = 400, 20_000, 5
    p = CNVMParameters(num_opinions=3, num_agents=N, r=R3, r_tilde=RT3, alpha=1.0)
    t, c = sample_stochastic_approximation(p, [0.2, 0.5, 0.3], 3.0, R, T, rng=np.random.default_rng(3))
    assert c.shape == (R, T, 3) and np.allclose(c.sum(-1), 1.0)
    np.testing.assert_allclose(c[:, 0], np.broadcast_to([0.2, 0.5, 0.3], (R, 3)))
    x0 = np.repeat(np.arange(3), [80, 200, 120])
    _, a = sample_many_runs(p, x0, 3.0, T, R, collective_variable=OpinionShares(3, normalize=True), rng=np.random.default_rng(4))
    se = np.sqrt(c.var(0) / R + a.var(0) / R)
    z = np.abs(c.mean(0) - a.mean(0))[1:] / se[1:]
    assert z.max() < 5, z.max()
    two = sample_stochastic_approximation(p, [[0.2, 0.5, 0.3], [1.0, 0.0, 0.0]], 1.0, 10, 3, seed=5)[1]
    assert two.shape == (2, 10, 3, 3)
