"""CPU: the oracle (oracle/sponet_oracle.c) against fixtures produced by the unmodified reference.

This is what pins the oracle (tests/golden/make_golden.py generated the fixtures from /root/reference).
Everything is compared bit for bit: event times (float64) and agent states.
"""
import numpy as np
import pytest
from numpy.random import PCG64, Generator

from conftest import golden_names, load_golden, raw_stream

STREAM = 4_000_000


def _t_eval(g):
    te = g["t_eval"]
    if te.ndim == 0:
        v = int(te)
        return None if v < 0 else np.linspace(0, float(g["t_max"]), v)
    return te.astype(float)


@pytest.mark.parametrize("name", golden_names("cnvm_"))
def test_cnvm_fixture(oracle, name):
    g = load_golden(name)
    M = int(g["num_opinions"])
    s = oracle.Stream(raw_stream(str(g["bitgen"]), int(g["seed"]), STREAM))
    te = _t_eval(g)
    complete = bool(g["complete"])
    args = (float(g["r_imit"]), float(g["r_noise"]), g["prob_imit"], g["prob_noise"])
    if te is None:
        rp, col = (None, None) if complete else (g["row_ptr"], g["col"])
        t, x, _ = oracle.cnvm_all(g["x_init"], float(g["t_max"]), M, rp, col, *args, g["degree_alpha"], s)
    elif complete:
        t, x, _ = oracle.cnvm_complete_teval(g["x_init"], te, M, *args, g["degree_alpha"][0], s)
    else:
        t, x, _ = oracle.cnvm_teval(g["x_init"], te, M, g["row_ptr"], g["col"], *args, g["degree_alpha"], s)
    assert x.dtype == g["x_traj"].dtype
    np.testing.assert_array_equal(t, g["t_traj"])
    np.testing.assert_array_equal(x, g["x_traj"])


@pytest.mark.parametrize("name", golden_names("cntm_"))
def test_cntm_fixture(oracle, name):
    g = load_golden(name)
    s = oracle.Stream(raw_stream(str(g["bitgen"]), int(g["seed"]), STREAM))
    te = _t_eval(g)
    args = (float(g["next_event_rate"]), float(g["noise_prob"]), float(g["threshold_01"]), float(g["threshold_10"]))
    if te is None:
        t, x, _ = oracle.cntm_all(g["x_init"], float(g["t_max"]), *args, g["row_ptr"], g["col"], s)
    else:
        t, x, _ = oracle.cntm_teval(g["x_init"], te, *args, g["row_ptr"], g["col"], s)
    np.testing.assert_array_equal(t, g["t_traj"])
    np.testing.assert_array_equal(x, g["x_traj"])


def _cnvm_rates(r, rt):
    M = r.shape[0]
    r_imit, r_noise = r.max(), rt.max() * M
    return r_imit, r_noise, r / r_imit, rt * M / r_noise


@pytest.mark.parametrize("name", golden_names("smr_"))
def test_sample_many_runs_stream_discipline(oracle, name):
    """Serial: ONE stream threads through all runs; n_jobs=2: rng.spawn streams per chunk
    (sponet/multiprocessing.py:77-108)."""
    g = load_golden(name)
    states, R, T = g["x_init"], int(g["num_runs"]), int(g["num_t"])
    S, N = states.shape
    te = np.linspace(0, float(g["t_max"]), T)
    r_imit, r_noise, pi, pn = _cnvm_rates(g["r"], g["r_tilde"])
    da = np.ones(N)  # alpha = 1
    rng = Generator(PCG64(int(g["seed"])))
    n_jobs = int(g["n_jobs"])
    if n_jobs == 0:
        blocks = [(rng, range(S), range(R))]
    else:
        rngs = rng.spawn(n_jobs)
        if R >= S:
            sizes = [R // n_jobs + (1 if w < R % n_jobs else 0) for w in range(n_jobs)]
            b = np.concatenate(([0], np.cumsum(sizes)))
            blocks = [(rngs[w], range(S), range(b[w], b[w + 1])) for w in range(n_jobs)]
        else:
            chunks = np.array_split(np.arange(S), n_jobs)
            blocks = [(rngs[w], chunks[w], range(R)) for w in range(n_jobs)]
    full = np.zeros((S, R, T, N), dtype=np.uint8)
    for gen, js, runs in blocks:
        s = oracle.Stream(gen.bit_generator.random_raw(STREAM))
        for j in js:
            for i in runs:
                _, x, _ = oracle.cnvm_teval(states[j], te, 3, g["row_ptr"], g["col"], r_imit, r_noise, pi, pn, da, s)
                full[j, i] = x
    expect = g["x_out"]
    if expect.shape[-1] == N:
        np.testing.assert_array_equal(full, expect)
    else:  # OpinionShares(3, normalize=True)
        cv = np.stack([oracle.opinion_shares(full[j, i], 3) / N for j in range(S) for i in range(R)]).reshape(S, R, T, 3)
        np.testing.assert_array_equal(cv, expect)


def test_cv_known_answers(oracle):
    g = load_golden("cv_known_answers")
    x, w, deg = g["x"].astype(np.uint8), g["weights"], g["degrees"].astype(float)
    np.testing.assert_array_equal(oracle.opinion_shares(x, 3), g["shares"])
    np.testing.assert_array_equal(oracle.opinion_shares(x, 3) / x.shape[1], g["shares_norm"])
    np.testing.assert_array_equal(oracle.opinion_shares(x, 3, w), g["shares_w"])
    np.testing.assert_array_equal(oracle.opinion_shares(x, 3, deg), g["shares_deg"])


def test_edge_cv_known_answers(oracle):
    """Interfaces / Propensities restatements against values computed by the reference's own classes
    (tests/golden/make_cv2_golden.py); the GPU tests use them to check the in-kernel evaluation."""
    g = load_golden("cv2_known_answers")
    rp, col, x = g["row_ptr"], g["col"], g["x"]
    np.testing.assert_array_equal(oracle.interfaces(x, rp, col), g["interfaces"])
    np.testing.assert_array_equal(oracle.interfaces(x, rp, col) / int(g["num_edges"]), g["interfaces_norm"])
    deg = np.diff(rp)
    for alpha in (1.0, 0.5, 0.0):
        out = oracle.propensities(x, rp, col, deg ** alpha, g["r"], g["r_tilde"])
        np.testing.assert_array_equal(out, g[f"prop_net_a{alpha}"])  # same sequential f64 sum: bit-exact
        np.testing.assert_array_equal(out / x.shape[1], g[f"prop_net_norm_a{alpha}"])


def test_pcg64_stream_matches_numpy(oracle):
    bg = PCG64(11)
    s = oracle.Stream(pcg64_state=bg.state)
    assert [s.next_u64() for _ in range(64)] == [int(v) for v in bg.random_raw(64)]


def test_philox4x32_known_answers(oracle):
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, out in kat:
        assert tuple(int(v) for v in oracle.philox4x32_10(ctr, key)) == out


def test_native_poisson_moments(oracle):
    """The interval-count sampler of native mode is Poisson: mean/variance within 5 sigma."""
    for lam in (0.7, 4.0, 37.5, 2600.0, 2.6e5):
        k = np.array([oracle.native_poisson(123, c, 1, lam) for c in range(20000)], dtype=float)
        n = k.size
        assert abs(k.mean() - lam) < 5 * np.sqrt(lam / n)
        assert abs(k.var() - lam) < 5 * lam * np.sqrt(2.0 / n) + 5 * np.sqrt(lam / n)
