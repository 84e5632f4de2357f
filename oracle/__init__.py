"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/sponet_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  ``sponet_b200`` (the product) never does.

Each function names the reference loop it restates; see the C file for line-level citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsponet_oracle.so")
_lib = None

_p = np.ctypeslib.ndpointer
_f64 = _p(np.float64, flags="C_CONTIGUOUS")
_i32 = _p(np.int32, flags="C_CONTIGUOUS")
_i64 = _p(np.int64, flags="C_CONTIGUOUS")
_u8 = _p(np.uint8, flags="C_CONTIGUOUS")
_u16 = _p(np.uint16, flags="C_CONTIGUOUS")
_u32 = _p(np.uint32, flags="C_CONTIGUOUS")
_u64 = _p(np.uint64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (plain gcc, no external libraries)."""
    srcs = [os.path.join(_HERE, f) for f in ("sponet_oracle.c", "oracle_loops.inc", "oracle_native.inc")]
    stale = not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_stream_from_buffer.restype = C.c_void_p
        _lib.oracle_stream_from_buffer.argtypes = [C.c_void_p, C.c_int64]
        _lib.oracle_stream_pcg64.restype = C.c_void_p
        _lib.oracle_stream_pcg64.argtypes = [C.c_uint64] * 4
        _lib.oracle_stream_pos.restype = C.c_int64
        _lib.oracle_stream_pos.argtypes = [C.c_void_p]
        _lib.oracle_stream_overflow.restype = C.c_int
        _lib.oracle_stream_overflow.argtypes = [C.c_void_p]
        _lib.oracle_stream_free.argtypes = [C.c_void_p]
        _lib.oracle_stream_next_u64.restype = C.c_uint64
        _lib.oracle_stream_next_u64.argtypes = [C.c_void_p]
        _lib.oracle_stream_next_exponential.restype = C.c_double
        _lib.oracle_stream_next_exponential.argtypes = [C.c_void_p]
        for name in (
            "oracle_cnvm_teval_u8", "oracle_cnvm_teval_u16", "oracle_cnvm_complete_teval_u8",
            "oracle_cnvm_complete_teval_u16", "oracle_cntm_teval", "oracle_cnvm_all_u8",
            "oracle_cnvm_all_u16", "oracle_cnvm_complete_all_u8", "oracle_cnvm_complete_all_u16",
            "oracle_cntm_all", "oracle_sample_many_runs_u8", "oracle_native_poisson",
            "oracle_cnvm_native", "oracle_cntm_native", "oracle_cnvm_complete_counts_native",
        ):
            getattr(_lib, name).restype = C.c_int64
    return _lib


class Stream:
    """An injected raw-uint64 stream (``BitGenerator.random_raw``) or a live numpy-compatible PCG64."""

    def __init__(self, raw: np.ndarray | None = None, pcg64_state: dict | None = None):
        L = lib()
        if raw is not None:
            self._raw = np.ascontiguousarray(raw, dtype=np.uint64)
            self._h = L.oracle_stream_from_buffer(self._raw.ctypes.data, self._raw.size)
        else:
            st, inc = pcg64_state["state"]["state"], pcg64_state["state"]["inc"]
            m = (1 << 64) - 1
            self._h = L.oracle_stream_pcg64(st >> 64, st & m, inc >> 64, inc & m)

    @property
    def pos(self) -> int:
        return lib().oracle_stream_pos(self._h)

    @property
    def overflow(self) -> bool:
        return bool(lib().oracle_stream_overflow(self._h))

    def next_u64(self) -> int:
        return lib().oracle_stream_next_u64(self._h)

    def next_exponential(self) -> float:
        return lib().oracle_stream_next_exponential(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and _lib is not None:
                _lib.oracle_stream_free(self._h)
        except Exception:  # interpreter shutdown
            pass
        self._h = None


def neighbor_list_to_csr(neighbor_list) -> tuple[np.ndarray, np.ndarray]:
    """Row order preserved: neighbour slot k is drawn as int(u*deg) (cnvm/model.py:286)."""
    deg = np.array([len(nb) for nb in neighbor_list], dtype=np.int64)
    row_ptr = np.zeros(len(neighbor_list) + 1, dtype=np.int32)
    np.cumsum(deg, out=row_ptr[1:])
    col = (
        np.concatenate([np.asarray(nb, dtype=np.int32) for nb in neighbor_list])
        if len(neighbor_list) and row_ptr[-1] > 0
        else np.zeros(0, dtype=np.int32)
    )
    return row_ptr, np.ascontiguousarray(col, dtype=np.int32)


def build_alias_table(weights) -> tuple[np.ndarray, np.ndarray]:
    """sponet/sampling.py:54-96."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    prob = np.empty_like(w)
    alias = np.empty(w.size, dtype=np.int32)
    f = lib().oracle_build_alias_table
    f.argtypes = [_f64, C.c_int64, _f64, _i32]
    f.restype = None
    f(w, w.size, prob, alias)
    return prob, alias


def _state(x, M):
    dt = np.uint8 if M <= 256 else np.uint16
    return np.array(x, dtype=dt), ("u8" if M <= 256 else "u16"), (_u8 if M <= 256 else _u16)


def _mats(M, prob_imit, prob_noise):
    pi = np.ascontiguousarray(prob_imit, dtype=np.float64).reshape(M, M)
    pn = np.ascontiguousarray(prob_noise, dtype=np.float64).reshape(M, M)
    return pi, pn


def cnvm_teval(x, t_eval, M, row_ptr, col, r_imit, r_noise, prob_imit, prob_noise, degree_alpha, stream):
    """sponet/cnvm/model.py:239-308.  Returns (t_traj, x_traj, iterations)."""
    x, suf, xt = _state(x, M)
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    T, n = t_eval.size, x.size
    pi, pn = _mats(M, prob_imit, prob_noise)
    t_out = np.empty(T)
    x_out = np.empty((T, n), dtype=x.dtype)
    f = getattr(lib(), f"oracle_cnvm_teval_{suf}")
    f.argtypes = [xt, C.c_int64, _f64, C.c_int64, C.c_int64, _i32, _i32, C.c_double, C.c_double,
                  _f64, _f64, _f64, C.c_void_p, _f64, xt]
    it = f(x, n, t_eval, T, M, row_ptr, col, r_imit, r_noise, pi, pn,
           np.ascontiguousarray(degree_alpha, dtype=np.float64), stream._h, t_out, x_out)
    if it < 0:
        raise RuntimeError("oracle: injected stream exhausted")
    return t_out, x_out, it


def cnvm_complete_teval(x, t_eval, M, r_imit, r_noise, prob_imit, prob_noise, max_degree_alpha, stream):
    """sponet/cnvm/model.py:365-433."""
    x, suf, xt = _state(x, M)
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    T, n = t_eval.size, x.size
    pi, pn = _mats(M, prob_imit, prob_noise)
    t_out = np.empty(T)
    x_out = np.empty((T, n), dtype=x.dtype)
    f = getattr(lib(), f"oracle_cnvm_complete_teval_{suf}")
    f.argtypes = [xt, C.c_int64, _f64, C.c_int64, C.c_int64, C.c_double, C.c_double, _f64, _f64,
                  C.c_double, C.c_void_p, _f64, xt]
    it = f(x, n, t_eval, T, M, r_imit, r_noise, pi, pn, float(max_degree_alpha), stream._h, t_out, x_out)
    if it < 0:
        raise RuntimeError("oracle: injected stream exhausted")
    return t_out, x_out, it


def cntm_teval(x, t_eval, next_event_rate, noise_prob, threshold_01, threshold_10, row_ptr, col, stream):
    """sponet/cntm/model.py:163-227."""
    x = np.array(x, dtype=np.uint8)
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    T, n = t_eval.size, x.size
    t_out = np.empty(T)
    x_out = np.empty((T, n), dtype=np.uint8)
    f = lib().oracle_cntm_teval
    f.argtypes = [_u8, C.c_int64, _f64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                  _i32, _i32, C.c_void_p, _f64, _u8]
    it = f(x, n, t_eval, T, next_event_rate, noise_prob, threshold_01, threshold_10, row_ptr, col,
           stream._h, t_out, x_out)
    if it < 0:
        raise RuntimeError("oracle: injected stream exhausted")
    return t_out, x_out, it


def _expand_log(x0, log_t, log_agent, log_op, cnt):
    """Turn the (t, agent, opinion) event log into the reference's x.copy()-per-change lists."""
    t_traj = np.concatenate([[0.0], log_t[:cnt]])
    x_traj = np.repeat(np.asarray(x0)[None, :], cnt + 1, axis=0)
    for i in range(cnt):
        x_traj[i + 1 :, log_agent[i]] = log_op[i]
    return t_traj, x_traj


def cnvm_all(x, t_max, M, row_ptr, col, r_imit, r_noise, prob_imit, prob_noise, degree_alpha, stream, cap=1 << 20):
    """sponet/cnvm/model.py:184-236 (row_ptr given) or :311-362 (row_ptr None)."""
    x, suf, xt = _state(x, M)
    x0 = x.copy()
    n = x.size
    pi, pn = _mats(M, prob_imit, prob_noise)
    log_t = np.empty(cap)
    log_a = np.empty(cap, dtype=np.int32)
    log_o = np.empty(cap, dtype=np.int32)
    iters = C.c_int64(0)
    da = np.ascontiguousarray(degree_alpha, dtype=np.float64)
    if row_ptr is not None:
        f = getattr(lib(), f"oracle_cnvm_all_{suf}")
        f.argtypes = [xt, C.c_int64, C.c_double, C.c_int64, _i32, _i32, C.c_double, C.c_double, _f64,
                      _f64, _f64, C.c_void_p, C.c_int64, _f64, _i32, _i32, C.POINTER(C.c_int64)]
        cnt = f(x, n, t_max, M, row_ptr, col, r_imit, r_noise, pi, pn, da, stream._h, cap, log_t,
                log_a, log_o, C.byref(iters))
    else:
        f = getattr(lib(), f"oracle_cnvm_complete_all_{suf}")
        f.argtypes = [xt, C.c_int64, C.c_double, C.c_int64, C.c_double, C.c_double, _f64, _f64,
                      C.c_double, C.c_void_p, C.c_int64, _f64, _i32, _i32, C.POINTER(C.c_int64)]
        cnt = f(x, n, t_max, M, r_imit, r_noise, pi, pn, float(da[0]), stream._h, cap, log_t, log_a,
                log_o, C.byref(iters))
    if cnt < 0:
        raise RuntimeError("oracle: stream exhausted" if cnt == -1 else "oracle: event log too small")
    t_traj, x_traj = _expand_log(x0, log_t, log_a, log_o, cnt)
    return t_traj, x_traj, iters.value


def cntm_all(x, t_max, next_event_rate, noise_prob, threshold_01, threshold_10, row_ptr, col, stream, cap=1 << 20):
    """sponet/cntm/model.py:111-160."""
    x = np.array(x, dtype=np.uint8)
    x0 = x.copy()
    log_t = np.empty(cap)
    log_a = np.empty(cap, dtype=np.int32)
    log_o = np.empty(cap, dtype=np.int32)
    iters = C.c_int64(0)
    f = lib().oracle_cntm_all
    f.argtypes = [_u8, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _i32,
                  _i32, C.c_void_p, C.c_int64, _f64, _i32, _i32, C.POINTER(C.c_int64)]
    cnt = f(x, x.size, t_max, next_event_rate, noise_prob, threshold_01, threshold_10, row_ptr, col,
            stream._h, cap, log_t, log_a, log_o, C.byref(iters))
    if cnt < 0:
        raise RuntimeError("oracle: stream exhausted" if cnt == -1 else "oracle: event log too small")
    t_traj, x_traj = _expand_log(x0, log_t, log_a, log_o, cnt)
    return t_traj, x_traj, iters.value


def opinion_shares(x, M, weights=None):
    """sponet/collective_variables.py:400-405 (np.bincount per snapshot)."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.uint8)
    out = np.empty((x.shape[0], M))
    f = lib().oracle_opinion_shares_u8
    f.argtypes = [_u8, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, _f64]
    f.restype = None
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    f(x, x.shape[0], x.shape[1], M, None if w is None else w.ctypes.data, out)
    return out


def interfaces(x, row_ptr, col):
    """sponet/collective_variables.py:293-301: number of edges {u, v} with x[u] != x[v], per snapshot
    (the reference loops over network.edges; here over the CSR entries with u < v)."""
    x = np.atleast_2d(np.asarray(x))
    rp = np.asarray(row_ptr, dtype=np.int64)
    src = np.repeat(np.arange(rp.size - 1), np.diff(rp))
    dst = np.asarray(col, dtype=np.int64)
    keep = src < dst
    return (x[:, src[keep]] != x[:, dst[keep]]).sum(axis=1).astype(np.float64)[:, None]


def propensities(x, row_ptr, col, degrees_alpha, r, r_tilde):
    """sponet/collective_variables.py:370-383 (_propensities_numba): sequential f64 sums over the agents j,
    out[i, m] += r[m, n] * #{neighbours of j with opinion n} / degrees_alpha[j] + r_tilde[m, n]."""
    x = np.atleast_2d(np.asarray(x))
    r, r_tilde = np.asarray(r, dtype=np.float64), np.asarray(r_tilde, dtype=np.float64)
    da = np.asarray(degrees_alpha, dtype=np.float64)
    out = np.zeros((x.shape[0], 2))
    for j in range(x.shape[1]):
        nbrs = col[row_ptr[j] : row_ptr[j + 1]]
        for i in range(x.shape[0]):
            m, n = (0, 1) if x[i, j] == 0 else (1, 0)
            count = np.float64(np.sum(x[i, nbrs] == n))
            out[i, m] += r[m, n] * count / da[j] + r_tilde[m, n]
    return out


def argmatch(x_ref, x):
    """sponet/utils.py:7-45."""
    x_ref = np.ascontiguousarray(x_ref, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x_ref.size, dtype=np.int64)
    f = lib().oracle_argmatch
    f.argtypes = [_f64, C.c_int64, _f64, C.c_int64, _i64]
    f.restype = None
    f(x_ref, x_ref.size, x, x.size, out)
    return out


def sample_many_runs(model, x_init, R, t_eval, M, row_ptr=None, col=None, r_imit=0.0, r_noise=0.0,
                     prob_imit=None, prob_noise=None, degree_alpha=None, next_event_rate=0.0,
                     noise_prob=0.0, threshold_01=0.0, threshold_10=0.0, seed=1, num_threads=1):
    """CPU baseline driver (multiprocessing.py:134-187 with OpinionShares), pthreads over runs.

    model: "cnvm" | "cnvm_complete" | "cntm".  Returns (counts [R,T,M] f64, total loop iterations).
    """
    code = {"cnvm": 0, "cnvm_complete": 1, "cntm": 2}[model]
    x_init = np.ascontiguousarray(x_init, dtype=np.uint8)
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    n, T = x_init.size, t_eval.size
    z32 = np.zeros(1, dtype=np.int32)
    zf = np.zeros(max(M * M, 1))
    row_ptr = z32 if row_ptr is None else row_ptr
    col = z32 if col is None else col
    pi = zf if prob_imit is None else np.ascontiguousarray(prob_imit, dtype=np.float64)
    pn = zf if prob_noise is None else np.ascontiguousarray(prob_noise, dtype=np.float64)
    da = zf if degree_alpha is None else np.ascontiguousarray(degree_alpha, dtype=np.float64)
    out = np.empty((R, T, M))
    f = lib().oracle_sample_many_runs_u8
    f.argtypes = [C.c_int, _u8, C.c_int64, C.c_int64, _f64, C.c_int64, C.c_int64, _i32, _i32,
                  C.c_double, C.c_double, _f64, _f64, _f64, C.c_double, C.c_double, C.c_double,
                  C.c_double, C.c_uint64, C.c_int, _f64]
    total = f(code, x_init, n, R, t_eval, T, M, row_ptr, col, r_imit, r_noise, pi, pn, da,
              next_event_rate, noise_prob, threshold_01, threshold_10, seed, num_threads, out)
    return out, total


# ---------------------------------------------------------------------------------------------
# native-RNG mode (sequential statement)
# ---------------------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    f = lib().oracle_philox4x32_10
    f.argtypes = [_u32, _u32, _u32]
    f.restype = None
    f(c, k, out)
    return out


def prob_to_thr(p) -> np.ndarray:
    """p -> floor(p * 2^32), saturating at 0xFFFFFFFF (which the compare treats as 'always')."""
    p = np.asarray(p, dtype=np.float64)
    t = np.floor(np.clip(p, 0.0, 1.0) * 4294967296.0)
    return np.minimum(t, 4294967295.0).astype(np.uint32)


def native_event_counts(seed, chain0, num_chains, t_eval, next_event_rate):
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    out = np.empty((num_chains, max(t_eval.size - 1, 0)), dtype=np.int64)
    f = lib().oracle_native_event_counts
    f.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, _f64, C.c_int64, C.c_double, _i64]
    f.restype = None
    f(seed, chain0, num_chains, t_eval, t_eval.size, next_event_rate, out)
    return out


def native_poisson(seed, chain, interval, lam):
    f = lib().oracle_native_poisson
    f.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double]
    return f(seed, chain, interval, lam)


def build_neighbor_table(row_ptr, col):
    """Independent numpy statement of the neighbour sampler table (include/sponet_b200.h,
    sponet_graph_neighbor_table).  Returns (table uint64 [n << lg], lg) or (None, -1) when the graph is
    not eligible (degree 0 or > 64, more than 2^20 agents, table above 64 MiB)."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    deg = np.diff(row_ptr)
    n = deg.size
    if deg.min() < 1 or deg.max() > 64 or n > (1 << 20):
        return None, -1
    lg = 1
    while (1 << lg) < deg.max():
        lg += 1
    L = 1 << lg
    if n * L * 8 > (64 << 20):
        return None, -1
    j = np.arange(L, dtype=np.int64)[None, :]
    d = deg[:, None]
    k = j * d // L
    num = (k + 1) * L - j * d
    cont = num < d
    base = row_ptr[:-1, None]
    colv = np.asarray(col, dtype=np.int64)
    id0 = colv[base + k]
    id1 = np.where(cont, colv[np.minimum(base + k + 1, colv.size - 1)], id0)
    thr = np.where(cont, (num << 24) // d, 0xFFFFFF)
    w0 = id0 | ((thr >> 12) << 20)
    w1 = id1 | ((thr & 0xFFF) << 20)
    return (w0 | (w1 << 32)).astype(np.uint64).ravel(), lg


def cnvm_native(x, M, row_ptr, col, noise_thr, thr_imit, thr_noise, alias_thr, alias, counts, seed, chain,
                want_x=True, nbtab=None, nbtab_log2=0):
    x = np.array(x, dtype=np.uint8)
    n = x.size
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    T = counts.size + 1
    out_x = np.empty((T, n), dtype=np.uint8) if want_x else None
    out_c = np.empty((T, M), dtype=np.int64)
    acc = C.c_int64(0)
    f = lib().oracle_cnvm_native
    f.argtypes = [_u8, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, _u32, _u32,
                  C.c_void_p, C.c_void_p, _i64, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p, _i64,
                  C.POINTER(C.c_int64)]
    if nbtab is not None:
        nbtab = np.ascontiguousarray(nbtab, dtype=np.uint64)
    ti = np.ascontiguousarray(thr_imit, dtype=np.uint32)
    tn = np.ascontiguousarray(thr_noise, dtype=np.uint32)
    keep = [row_ptr, col, alias_thr, alias, nbtab]
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    ev = f(x, n, M, ptr(row_ptr), ptr(col), ptr(nbtab), int(nbtab_log2), int(noise_thr), ti, tn, ptr(alias_thr), ptr(alias), counts,
           T, seed, chain, ptr(out_x), out_c, C.byref(acc))
    del keep
    return out_x, out_c, ev, acc.value


def cnvm_complete_counts_native(c0, n, M, noise_thr, thr_imit, thr_noise, counts, seed, chain):
    c = np.array(c0, dtype=np.int64)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    T = counts.size + 1
    out_c = np.empty((T, M), dtype=np.int64)
    f = lib().oracle_cnvm_complete_counts_native
    f.argtypes = [_i64, C.c_int64, C.c_int64, C.c_uint32, _u32, _u32, _i64, C.c_int64, C.c_uint64,
                  C.c_uint64, _i64]
    ev = f(c, n, M, int(noise_thr), np.ascontiguousarray(thr_imit, dtype=np.uint32),
           np.ascontiguousarray(thr_noise, dtype=np.uint32), counts, T, seed, chain, out_c)
    return out_c, ev


def cntm_native(x, row_ptr, col, noise_thr, threshold_01, threshold_10, counts, seed, chain, want_x=True):
    x = np.array(x, dtype=np.uint8)
    n = x.size
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    T = counts.size + 1
    out_x = np.empty((T, n), dtype=np.uint8) if want_x else None
    out_c = np.empty((T, 2), dtype=np.int64)
    acc = C.c_int64(0)
    f = lib().oracle_cntm_native
    f.argtypes = [_u8, C.c_int64, _i32, _i32, C.c_uint32, C.c_double, C.c_double, _i64, C.c_int64,
                  C.c_uint64, C.c_uint64, C.c_void_p, _i64, C.POINTER(C.c_int64)]
    ev = f(x, n, row_ptr, col, int(noise_thr), threshold_01, threshold_10, counts, T, seed, chain,
           None if out_x is None else out_x.ctypes.data, out_c, C.byref(acc))
    return out_x, out_c, ev, acc.value


def sample_states(n, M, counts, S, seed):
    """oracle_sample_states: counts None (uniform opinions) or int64 [S, M] / [1, M] (shared)."""
    out = np.empty((S, n), dtype=np.uint8)
    f = lib().oracle_sample_states
    f.restype = None
    f.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_uint64, _u8]
    if counts is None:
        f(n, M, None, 0, S, seed, out)
    else:
        c = np.ascontiguousarray(np.atleast_2d(counts), dtype=np.int64)
        f(n, M, c.ctypes.data, int(c.shape[0] == 1), S, seed, out)
    return out


def er_graph(n, p, seed, c, force_no_isolates=False, max_attempts=64):
    """oracle_er_graph: graph index c of the device-side Erdos-Renyi generator as (row_ptr, col, attempts)."""
    f = lib().oracle_er_graph
    f.restype = C.c_int64
    f.argtypes = [C.c_int64, C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.c_int32, _i32, _i32, C.c_int64,
                  C.POINTER(C.c_int32)]
    row_ptr = np.empty(n + 1, dtype=np.int32)
    cap = int(max(1024, 2 * (n * (n - 1) * p / 2 + 8 * np.sqrt(n * n * p) + 2 * n)))
    cap = min(cap, n * (n - 1) + 2 * n)
    col = np.empty(cap, dtype=np.int32)
    used = C.c_int32(0)
    nnz = f(n, float(p), int(bool(force_no_isolates)), seed, c, max_attempts, row_ptr, col, cap, C.byref(used))
    if nnz == -1:
        raise RuntimeError("Timeout. Could not generate a graph without isolated vertices.")
    if nnz < 0:
        raise RuntimeError("oracle_er_graph: column buffer too small")
    return row_ptr, col[:nnz].copy(), used.value


def ba_graph(n, m, seed, c):
    """oracle_ba_graph: graph index c of the device-side Barabasi-Albert generator as (row_ptr, col)."""
    f = lib().oracle_ba_graph
    f.restype = C.c_int64
    f.argtypes = [C.c_int64, C.c_int32, C.c_uint64, C.c_uint64, _i32, _i32]
    row_ptr = np.empty(n + 1, dtype=np.int32)
    col = np.empty(2 * m * (n - m), dtype=np.int32)
    nnz = f(n, m, seed, c, row_ptr, col)
    assert nnz == col.size
    return row_ptr, col


def ws_row_capacity(k: int) -> int:
    """Adjacency slots per vertex of the Watts-Strogatz generators (device and oracle): k lattice neighbours + rewired-in edges."""
    return int(k) + 32


def ws_graph(n, k, p, seed, c, tries=100):
    """oracle_ws_graph: graph index c of the device-side Watts-Strogatz generator as (row_ptr, col, attempts)."""
    f = lib().oracle_ws_graph
    f.restype = C.c_int64
    f.argtypes = [C.c_int64, C.c_int32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, _i32, _i32, C.POINTER(C.c_int32)]
    row_ptr = np.empty(n + 1, dtype=np.int32)
    col = np.empty(n * (k // 2) * 2, dtype=np.int32)
    used = C.c_int32(0)
    nnz = f(n, k, int(prob_to_thr(np.array([p]))[0]), seed, c, tries, ws_row_capacity(k), row_ptr, col, C.byref(used))
    if nnz == -1:
        raise RuntimeError("Maximum number of tries exceeded")
    if nnz < 0:
        raise RuntimeError("oracle_ws_graph: a vertex exceeds the row capacity")
    assert nnz == col.size
    return row_ptr, col, used.value


def regular_graph(n, d, seed, c, tries=64):
    """oracle_regular_graph: graph index c of the device-side random regular generator as (row_ptr, col, attempts)."""
    f = lib().oracle_regular_graph
    f.restype = C.c_int64
    f.argtypes = [C.c_int64, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, _i32, _i32, C.POINTER(C.c_int32)]
    row_ptr = np.empty(n + 1, dtype=np.int32)
    col = np.empty(n * d, dtype=np.int32)
    used = C.c_int32(0)
    nnz = f(n, d, seed, c, tries, row_ptr, col, C.byref(used))
    if nnz < 0:
        raise RuntimeError("oracle_regular_graph: no attempt produced a regular graph")
    return row_ptr, col, used.value
