"""Native mode's snapshot rule against the reference's, on the reference's own trajectories (CPU).

Native mode stores the chain state AT ``t_eval[k]``: the state after
``1 + #{arrivals before t_eval[k]}`` loop iterations.  The reference stores the state at the event
nearest to ``t_eval[k]`` (sponet/utils.py:96-135): either the same state, or that state with the LAST
ACCEPTED change undone, stamped with that change's time.  This file pins the two claims the native
mode rests on:

  1. iteration count: when the reference stores output k it has executed exactly
     ``1 + #{i : t_i < t_eval[k]}`` loop iterations, the t_i being the partial sums of its
     exponential draws -- so K_1 - 1, K_2, ... are the Poisson(lambda * dt_k) interval counts the
     native scheduler draws;
  2. bias bound: the reference snapshot and the at-time state differ in AT MOST ONE agent (the last
     accepted one), i.e. every opinion count differs by at most 1 and every share by at most 1/N,
     and the stamped time by at most one inter-arrival gap on either side.

The trajectory is produced by a line-by-line Python restatement of ``_simulate_teval``
(sponet/cnvm/model.py:239-308) that ALSO records the un-reverted state; its reference outputs are
checked against the C oracle (which the reference-generated fixtures pin) on the same injected stream.
"""
import numpy as np
import pytest
from numpy.random import PCG64

R3 = np.array([[0, 1, 2], [1, 0, 1], [2, 0, 0]], dtype=float)
RT3 = np.array([[0, 0.2, 0.1], [0, 0, 0.1], [0.1, 0.2, 0]], dtype=float)


def _teval_with_at_time_state(oracle, x, t_eval, M, nl, r_imit, r_noise, prob_imit, prob_noise, degree_alpha, stream):
    """cnvm/model.py:239-308 in Python.  Returns (t_traj, x_traj, x_at, iters, arrivals): x_at[k] is the
    loop's state when output k is stored (before the possible undo), iters[k] the iterations executed until
    then, arrivals the event times t_1, t_2, ... (partial sums of the exponentials)."""
    u = lambda: (stream.next_u64() >> 11) * 2.0**-53  # noqa: E731  numba's next_double
    n = x.size
    total = 0.0
    for d in degree_alpha:  # numba's np.sum is a sequential loop
        total += float(d)
    next_event_rate = 1.0 / (r_imit * total + r_noise * n)
    noise_probability = r_noise * n * next_event_rate
    prob_table, alias_table = oracle.build_alias_table(degree_alpha)
    x = x.copy()
    t = t_eval[0]
    t_traj, x_traj, x_at, iters, arrivals = [t], [x.copy()], [x.copy()], [0], []
    previous_agent, previous_opinion, previous_t = 0, x[0], t
    k, it = 1, 0
    while k < len(t_eval):
        it += 1
        if u() < noise_probability:
            agent = int(u() * n)
            new_opinion = int(u() * M)
            accept = u() < prob_noise[x[agent], new_opinion]
        else:
            r = u() * n  # sampling.py:117-122: one uniform -> index and fraction
            i = int(r)
            agent = i if (r - i) < prob_table[i] else int(alias_table[i])
            nb = nl[agent]
            new_opinion = x[nb[int(u() * len(nb))]]
            accept = u() < prob_imit[x[agent], new_opinion]
        if accept:
            previous_t, previous_agent, previous_opinion = t, agent, x[agent]
            x[agent] = new_opinion
        t += next_event_rate * stream.next_exponential()
        arrivals.append(t)
        if t >= t_eval[k]:
            x_at.append(x.copy())
            iters.append(it)
            xs = x.copy()
            if t - t_eval[k] <= abs(t_eval[k] - previous_t):
                t_traj.append(t)
            else:
                xs[previous_agent] = previous_opinion
                t_traj.append(previous_t)
            x_traj.append(xs)
            k += 1
    return np.array(t_traj), np.array(x_traj), np.array(x_at), np.array(iters), np.array(arrivals)


@pytest.mark.parametrize("seed,n,T,t_max", [(1, 60, 12, 3.0), (2, 45, 40, 1.0), (3, 80, 5, 6.0)])
def test_reference_snapshot_vs_state_at_t_eval(oracle, seed, n, T, t_max):
    import networkx as nx

    from sponet_b200 import CNVMParameters

    g = nx.fast_gnp_random_graph(n, 6 / (n - 1), seed=seed)
    for v in list(nx.isolates(g)):
        g.add_edge(v, (v + 1) % n)
    nl = [np.fromiter(g.neighbors(v), dtype=np.int32) for v in g.nodes()]
    p = CNVMParameters(num_opinions=3, network=nl, r=R3, r_tilde=RT3, alpha=0.5)
    da = np.array([len(nb) for nb in nl], dtype=float) ** 0.5  # d^(1 - alpha)
    x0 = np.random.default_rng(seed).integers(0, 3, n).astype(np.uint8)
    t_eval = np.linspace(0.0, t_max, T)
    raw = PCG64(seed).random_raw(200_000)
    rp, col = oracle.neighbor_list_to_csr(nl)

    t_ref, x_ref, it_ref = oracle.cnvm_teval(x0, t_eval, 3, rp, col, p.r_imit, p.r_noise, p.prob_imit, p.prob_noise, da,
                                             oracle.Stream(raw))
    t_py, x_py, x_at, iters, arrivals = _teval_with_at_time_state(oracle, x0, t_eval, 3, nl, p.r_imit, p.r_noise,
                                                                  p.prob_imit, p.prob_noise, da, oracle.Stream(raw))
    # the Python restatement IS the reference loop (the C oracle is pinned by reference-generated fixtures)
    np.testing.assert_array_equal(t_py, t_ref)
    np.testing.assert_array_equal(x_py, x_ref)
    assert iters[-1] == it_ref

    reverted = 0
    for k in range(1, T):
        # 1. iteration count = 1 + arrivals strictly before t_eval[k]  (the native interval schedule)
        assert iters[k] == 1 + int(np.sum(arrivals < t_eval[k]))
        # 2. at most one agent differs, every count by at most 1, the time stamp within one gap
        diff = np.flatnonzero(x_ref[k] != x_at[k])
        assert diff.size <= 1
        reverted += diff.size
        c_ref, c_at = np.bincount(x_ref[k], minlength=3), np.bincount(x_at[k], minlength=3)
        assert np.abs(c_ref - c_at).max() <= 1
        after = arrivals[arrivals >= t_eval[k]][0]
        before = np.concatenate(([t_eval[0]], arrivals[arrivals < t_eval[k]]))
        assert before.min() <= t_ref[k] <= after
    # the undo branch is taken with probability p_acc / (1 + p_acc) per output: make sure the test saw it
    if T >= 12:
        assert reverted >= 1


def test_interval_counts_are_poisson(oracle):
    """Claim 1 in distribution: the native scheduler's counts per interval, K_1 - 1, K_2, ..., have the
    Poisson(lambda * dt) mean and variance (oracle statement of the in-kernel sampler, both of its branches)."""
    t_eval = np.array([0.0, 0.5, 0.50001, 2.0])
    rate = 1.0 / 4000.0  # next_event_rate: lambda = 4000 per unit time -> lam = 2000 (PTRS), 0.04 (product), 6000
    K = oracle.native_event_counts(123, 0, 20_000, t_eval, rate).astype(float)
    K[:, 0] -= 1
    lam = np.diff(t_eval) / rate
    for j in range(3):
        m, v, nn = K[:, j].mean(), K[:, j].var(), K.shape[0]
        assert abs(m - lam[j]) <= 5 * np.sqrt(lam[j] / nn), (j, m, lam[j])
        assert abs(v - lam[j]) <= 5 * np.sqrt((lam[j] + 2 * lam[j] ** 2) / nn) + 1e-9, (j, v, lam[j])
