"""Thin host layer between the reference-shaped Python API and the C ABI (include/sponet_b200.h).

Nothing here computes: it validates, marshals numpy / torch buffers into the ABI structs, sizes the
injected streams for replay mode and retries when a stream was too short.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
from numpy.random import Generator

from . import _lib
from .utils import neighbor_list_to_csr

_DEFAULT_MODE = os.environ.get("SPONET_B200_MODE", "native")


def set_default_mode(mode: str) -> None:
    """``"native"`` (Philox, production path) or ``"replay"`` (bit-exact re-execution of the reference)."""
    global _DEFAULT_MODE
    if mode not in ("native", "replay"):
        raise ValueError("mode must be 'native' or 'replay'")
    _DEFAULT_MODE = mode


def resolve_mode(mode: str | None) -> str:
    mode = _DEFAULT_MODE if mode is None else mode
    if mode not in ("native", "replay"):
        raise ValueError("mode must be 'native' or 'replay'")
    return mode


def current_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


class GraphHandle:
    """Device-resident CSR of one network on one GPU (sponet_graph_create / _destroy)."""

    def __init__(self, neighbor_list):
        _lib.require_device()
        self.row_ptr, self.col = neighbor_list_to_csr(neighbor_list)
        self.num_agents = len(neighbor_list)
        self.degrees = np.diff(self.row_ptr)
        h = C.c_void_p()
        _lib.check(_lib.lib().sponet_graph_create(self.num_agents, _lib.ptr(self.row_ptr), _lib.ptr(self.col), C.byref(h)))
        self._h = h
        self.device = current_device()

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib().sponet_graph_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close


class GraphBatch:
    """A batch of Erdos-Renyi graphs generated on the device (sponet_graph_batch_create_er), one per chain of the native
    call over the same (states, run range): chain (state j, run i) runs on graph index j * num_runs_total + i, which is a
    function of (seed, index) only.  Stands in for the reference's per-run ``update_network()`` (cnvm/model.py:95-96)."""

    def __init__(self, num_agents, p, seed, num_states, num_runs_total, run_begin, run_end, force_no_isolates=False,
                 max_attempts=64, padded_rows=False, records=True):
        _lib.require_device()
        h = C.c_void_p()
        rc = _lib.lib().sponet_graph_batch_create_er(int(num_agents), float(p), int(bool(force_no_isolates)), int(seed),
                                                     int(num_states), int(num_runs_total), int(run_begin), int(run_end),
                                                     int(max_attempts), int(bool(padded_rows)) | 2 * int(bool(records)), C.byref(h))
        if rc == _lib.ERR_INVALID and _lib.last_error().startswith("Timeout"):
            raise RuntimeError(_lib.last_error())  # network_generator.py:75-78
        _lib.check(rc)
        self._adopt(h, num_agents, num_states, run_begin, run_end)

    def _adopt(self, h, num_agents, num_states, run_begin, run_end):
        self._h = h
        self.num_agents = int(num_agents)
        self.count = int(num_states) * (int(run_end) - int(run_begin))
        self.device = current_device()

    @classmethod
    def barabasi_albert(cls, num_agents, m, seed, num_states, num_runs_total, run_begin, run_end, padded_rows=False, records=True):
        """Barabasi-Albert graphs (sponet_graph_batch_create_ba; networkx.barabasi_albert_graph's process)."""
        _lib.require_device()
        h = C.c_void_p()
        _lib.check(_lib.lib().sponet_graph_batch_create_ba(int(num_agents), int(m), int(seed), int(num_states), int(num_runs_total),
                                                           int(run_begin), int(run_end),
                                                           int(bool(padded_rows)) | 2 * int(bool(records)), C.byref(h)))
        self = cls.__new__(cls)
        self._adopt(h, num_agents, num_states, run_begin, run_end)
        return self

    @property
    def handle(self):
        return self._h

    @classmethod
    def watts_strogatz(cls, num_agents, k, p, seed, num_states, num_runs_total, run_begin, run_end, tries=100, padded_rows=False,
                       records=True):
        """Connected Watts-Strogatz graphs (sponet_graph_batch_create_ws; networkx.connected_watts_strogatz_graph's process)."""
        _lib.require_device()
        h = C.c_void_p()
        rc = _lib.lib().sponet_graph_batch_create_ws(int(num_agents), int(k), float(p), int(seed), int(num_states),
                                                     int(num_runs_total), int(run_begin), int(run_end), int(tries),
                                                     int(bool(padded_rows)) | 2 * int(bool(records)), C.byref(h))
        if rc == _lib.ERR_INVALID and _lib.last_error().startswith("Maximum number of tries"):
            raise RuntimeError(_lib.last_error())  # networkx raises NetworkXError here
        _lib.check(rc)
        self = cls.__new__(cls)
        self._adopt(h, num_agents, num_states, run_begin, run_end)
        return self

    @classmethod
    def random_regular(cls, num_agents, d, seed, num_states, num_runs_total, run_begin, run_end, tries=1000, padded_rows=False,
                       records=True):
        """Random d-regular graphs (sponet_graph_batch_create_regular; networkx.random_regular_graph's pairing process)."""
        _lib.require_device()
        h = C.c_void_p()
        _lib.check(_lib.lib().sponet_graph_batch_create_regular(int(num_agents), int(d), int(seed), int(num_states),
                                                                int(num_runs_total), int(run_begin), int(run_end), int(tries),
                                                                int(bool(padded_rows)) | 2 * int(bool(records)), C.byref(h)))
        self = cls.__new__(cls)
        self._adopt(h, num_agents, num_states, run_begin, run_end)
        return self

    def set_weights(self, alpha: float) -> None:
        """Degree-weighted agent selection (alpha != 1): per-graph alias tables and weight sums on the device
        (sponet_graph_batch_set_weights); the weights are ``degree ** (1 - alpha)`` as cnvm/model.py:38 computes them."""
        dmax = self.info()["max_degree"]
        with np.errstate(divide="ignore"):
            w = np.ascontiguousarray(np.arange(dmax + 1, dtype=np.int64) ** (1 - alpha), dtype=np.float64)
        _lib.check(_lib.lib().sponet_graph_batch_set_weights(self._h, _lib.ptr(w), int(w.size)))

    def csr(self, slot: int):
        """Graph `slot` of the batch as host (row_ptr, col)."""
        row_ptr = np.empty(self.num_agents + 1, dtype=np.int32)
        L = _lib.lib()
        _lib.check(L.sponet_graph_batch_get(self._h, int(slot), _lib.ptr(row_ptr), None, 0))
        col = np.empty(max(int(row_ptr[-1]), 1), dtype=np.int32)
        _lib.check(L.sponet_graph_batch_get(self._h, int(slot), _lib.ptr(row_ptr), _lib.ptr(col), col.size))
        return row_ptr, col[: int(row_ptr[-1])]

    def info(self):
        n, nnz, dmin, dmax, dev = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_int()
        _lib.check(_lib.lib().sponet_graph_info(self._h, C.byref(n), C.byref(nnz), C.byref(dmin), C.byref(dmax), C.byref(dev)))
        return {"num_agents": n.value, "nnz": nnz.value, "min_degree": dmin.value, "max_degree": dmax.value, "device": dev.value}

    def close(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib().sponet_graph_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close


def trim_graph_batch_memory() -> None:
    """Returns the device memory kept from destroyed graph batches (sponet_graph_batch_trim) to the driver."""
    _lib.check(_lib.lib().sponet_graph_batch_trim())


class GraphCache:
    """One GraphHandle per (network identity, device); rebuilt when the adjacency object changes."""

    def __init__(self):
        self._obj = None
        self._handles: dict[int, GraphHandle] = {}

    def get(self, neighbor_list) -> GraphHandle:
        if neighbor_list is not self._obj:
            self.clear()
            self._obj = neighbor_list
        dev = current_device()
        if dev not in self._handles:
            self._handles[dev] = GraphHandle(neighbor_list)
        return self._handles[dev]

    def clear(self):
        for h in self._handles.values():
            h.close()
        self._handles = {}
        self._obj = None

    def detach(self) -> list:
        """Hands the current handles to the caller (who closes them once the kernels using them are done):
        per-run networks keep several graphs alive on different streams while the model moves on."""
        hs = list(self._handles.values())
        self._handles = {}
        self._obj = None
        return hs


# ---------------------------------------------------------------------------------------------
# struct marshalling
# ---------------------------------------------------------------------------------------------
def cnvm_struct(params, degree_alpha):
    """sponet_cnvm_params + the arrays it points to (keep both alive for the duration of the call)."""
    pi = np.ascontiguousarray(params.prob_imit, dtype=np.float64)
    pn = np.ascontiguousarray(params.prob_noise, dtype=np.float64)
    da = np.ascontiguousarray(degree_alpha, dtype=np.float64)
    s = _lib.sponet_cnvm_params(
        int(params.num_opinions), float(params.r_imit), float(params.r_noise), _lib.ptr(pi), _lib.ptr(pn), _lib.ptr(da)
    )
    return s, (pi, pn, da)


def cntm_struct(next_event_rate, noise_prob, params):
    return _lib.sponet_cntm_params(
        float(next_event_rate), float(noise_prob), float(params.threshold_01), float(params.threshold_10)
    )


def cnvm_event_rates(params, degree_alpha, complete: bool) -> tuple[float, float]:
    """(next_event_rate, noise_probability) with numba's sequential sum (cnvm/model.py:257-258, :383-384)."""
    ner, npr = C.c_double(), C.c_double()
    da = np.ascontiguousarray(degree_alpha, dtype=np.float64)
    _lib.check(
        _lib.lib().sponet_cnvm_event_rates(
            None if complete else _lib.ptr(da), int(params.num_agents), float(da[0]), float(params.r_imit),
            float(params.r_noise), C.byref(ner), C.byref(npr),
        )
    )
    return ner.value, npr.value


class KernelCV:
    """What the simulation kernels need to evaluate a collective variable on the fly (sponet_shares_cv).

    kind 0 (OpinionShares family): out[d] = count[idx[d]] / divisor(s), count = integer table
    [class * M + opinion] kept up to date on every accepted event.
    kind 1 / 2 (Interfaces / Propensities, two opinions): a scan of ``graph`` against the on-chip chain
    state at every output time; ``divisor`` is the normalisation."""

    SHARES, INTERFACES, PROPENSITIES = 0, 1, 2

    def __init__(self, idx, divisor=1.0, divisors=None, num_classes=1, agent_class=None, agent_weight=None,
                 kind=0, graph=None, degree_alpha=None, rates=None):
        self.idx = np.ascontiguousarray(idx, dtype=np.int32)
        self.divisor = float(divisor)
        self.divisors = None if divisors is None else np.ascontiguousarray(divisors, dtype=np.float64)
        self.num_classes = int(num_classes)
        self.agent_class = None if agent_class is None else np.ascontiguousarray(agent_class, dtype=np.int32)
        self.agent_weight = None if agent_weight is None else np.ascontiguousarray(agent_weight, dtype=np.int32)
        self.kind = int(kind)
        self.graph = graph  # GraphHandle of the CV's own network (kinds 1, 2)
        self.degree_alpha = None if degree_alpha is None else np.ascontiguousarray(degree_alpha, dtype=np.float64)
        self.rates = (0.0, 0.0, 0.0, 0.0) if rates is None else tuple(float(v) for v in rates)

    @property
    def grouped(self) -> bool:
        return self.agent_class is not None or self.agent_weight is not None

    @property
    def edge(self) -> bool:
        """Interfaces / Propensities: needs the agent states (never the count-only kernel)."""
        return self.kind != 0

    @property
    def integer_valued(self) -> bool:
        """Whether the kernels can accumulate exact integer moment sums of this CV."""
        return self.kind != self.PROPENSITIES

    @property
    def dimension(self) -> int:
        return {0: int(self.idx.size), 1: 1, 2: 2}[self.kind]

    def divisor_vector(self) -> np.ndarray:
        return self.divisors if self.divisors is not None else np.full(self.dimension, self.divisor)


def as_kernel_cv(spec) -> "KernelCV | None":
    if spec is None or isinstance(spec, KernelCV):
        return spec
    idx, divisor = spec  # legacy (idx, divisor) tuples
    return KernelCV(idx, divisor)


def cv_struct(spec):
    spec = as_kernel_cv(spec)
    if spec is None:
        return None, None
    s = _lib.sponet_shares_cv(
        spec.dimension, _lib.ptr(spec.idx), spec.divisor, spec.num_classes, _lib.ptr(spec.agent_class),
        _lib.ptr(spec.agent_weight), _lib.ptr(spec.divisors), spec.kind,
        None if spec.graph is None else spec.graph.handle, _lib.ptr(spec.degree_alpha),
        (_lib.C.c_double * 4)(*spec.rates),
    )
    return s, spec


def draw_seed(rng: Generator) -> int:
    """One uint64 from the caller's generator keys the Philox stream of a native-mode call."""
    return int(rng.integers(0, 2**64, dtype=np.uint64))


# ---------------------------------------------------------------------------------------------
# native mode
# ---------------------------------------------------------------------------------------------
def _alloc(shape, dtype, like):
    """Output buffer next to the inputs: torch CUDA tensor if ``like`` is one, numpy otherwise."""
    if like is not None and hasattr(like, "data_ptr") and like.is_cuda:
        import torch

        tdt = {np.uint8: torch.uint8, np.float64: torch.float64, np.int64: torch.int64, np.uint64: torch.uint64}[dtype]
        return torch.empty(shape, dtype=tdt, device=like.device)
    return np.empty(shape, dtype=dtype)


class Histogram:
    """Request for the in-kernel histogram of a CV over the runs (sponet_native_opts.hist_*): ``bins`` equal-width
    bins over the RAW integer CV values (counts before the divisor): bin = clamp((c - lo) // width, 0, bins - 1).
    ``out`` is the uint64 array [S, T, D, bins] the kernel ACCUMULATES into (numpy or CUDA tensor)."""

    def __init__(self, bins: int, lo: int, width: int, out):
        self.bins, self.lo, self.width, self.out = int(bins), int(lo), max(1, int(width)), out

    @staticmethod
    def for_counts(bins: int, max_count: int):
        """(lo, width) covering raw values 0 .. max_count with ``bins`` bins."""
        return 0, max(1, -(-(int(max_count) + 1) // int(bins)))

    def edges(self, divisor: float = 1.0) -> np.ndarray:
        """Bin edges in units of the CV (raw edges divided by the divisor); the last bin is closed on the right."""
        return (self.lo + self.width * np.arange(self.bins + 1)) / float(divisor)


def _native_opts(seed, num_runs_total, run_begin, run_end, event_counts, force_global_state, warps_per_chain,
                 final_state_only, hist, chain_offset=0):
    return _lib.sponet_native_opts(
        seed & (2**64 - 1), num_runs_total, run_begin, run_end, _lib.ptr(event_counts), int(force_global_state),
        int(warps_per_chain), int(final_state_only), 0 if hist is None else hist.bins, 0 if hist is None else hist.lo,
        1 if hist is None else hist.width, None if hist is None else _lib.ptr(hist.out), int(chain_offset),
    )


def run_native(
    kind: str,
    graph: GraphHandle | None,
    num_agents: int,
    model_struct,
    x_init,
    t_eval: np.ndarray,
    seed: int,
    num_runs_total: int,
    run_begin: int,
    run_end: int,
    cv_spec=None,
    want_x: bool = False,
    want_cv: bool = False,
    sums=None,
    event_counts=None,
    counters=None,
    force_global_state: bool = False,
    counts_init=None,
    warps_per_chain: int = 0,
    final_state_only: bool = False,
    hist: Histogram | None = None,
    chain_offset: int = 0,
):
    """One C-ABI call of the native ensemble path.

    kind: "cnvm" | "cnvm_counts" | "cntm".  ``x_init`` is (S, N) uint8, numpy or CUDA tensor; outputs
    are allocated on the same side.  ``sums`` = (sum, sumsq) int64 arrays [S, T, D] accumulated into;
    ``hist`` an in-kernel histogram request; ``final_state_only`` returns out_x as [S, nrun, N] (the state at
    the last output time).  Returns (out_x or None, out_cv or None).
    """
    L = _lib.lib()
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    T = t_eval.size
    nrun = run_end - run_begin
    like = counts_init if kind == "cnvm_counts" else x_init
    S = like.shape[0]
    cvs, cv_keep = cv_struct(cv_spec)
    D = 0 if cv_spec is None else cv_keep.dimension
    x_shape = (S, nrun, num_agents) if final_state_only else (S, nrun, T, num_agents)
    out_x = _alloc(x_shape, np.uint8, like) if want_x else None
    out_cv = _alloc((S, nrun, T, D), np.float64, like) if want_cv else None
    opts = _native_opts(seed, num_runs_total, run_begin, run_end, event_counts, force_global_state, warps_per_chain,
                        final_state_only, hist, chain_offset)
    sum_p = sumsq_p = None
    if sums is not None:
        sum_p, sumsq_p = _lib.ptr(sums[0]), _lib.ptr(sums[1])
    stream = _lib.current_stream_ptr()
    cvp = None if cvs is None else C.byref(cvs)
    if kind == "cnvm":
        rc = L.sponet_cnvm_run(
            graph.handle if graph is not None else None, num_agents, C.byref(model_struct), _lib.ptr(x_init), S,
            _lib.ptr(t_eval), T, C.byref(opts), cvp, _lib.ptr(out_x), _lib.ptr(out_cv), sum_p, sumsq_p,
            _lib.ptr(counters), stream,
        )
    elif kind == "cnvm_counts":
        rc = L.sponet_cnvm_run_complete_counts(
            num_agents, C.byref(model_struct), _lib.ptr(counts_init), S, _lib.ptr(t_eval), T, C.byref(opts), cvp,
            _lib.ptr(out_cv), sum_p, sumsq_p, _lib.ptr(counters), stream,
        )
    elif kind == "cntm":
        rc = L.sponet_cntm_run(
            graph.handle, C.byref(model_struct), _lib.ptr(x_init), S, _lib.ptr(t_eval), T, C.byref(opts), cvp,
            _lib.ptr(out_x), _lib.ptr(out_cv), sum_p, sumsq_p, _lib.ptr(counters), stream,
        )
    else:
        raise ValueError(kind)
    _lib.check(rc)
    return out_x, out_cv


def run_native_sweep(
    graph: GraphHandle | None,
    num_agents: int,
    model_structs: list,
    x_init,
    t_eval: np.ndarray,
    seed: int,
    num_runs_total: int,
    run_begin: int,
    run_end: int,
    cv_spec=None,
    want_cv: bool = False,
    sums=None,
    counters=None,
    warps_per_chain: int = 0,
    hist: Histogram | None = None,
    chain_offset: int = 0,
):
    """sponet_cnvm_run_sweep: P parameter sets (``model_structs``: sponet_cnvm_params sharing one degree_alpha) on
    one network in one launch; point p replaces the initial-state index in every output ([P, nrun, T, D],
    sums [P, T, D]).  ``x_init`` is (N,) / (1, N) (shared by all points) or (P, N)."""
    L = _lib.lib()
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    T, P = t_eval.size, len(model_structs)
    nrun = run_end - run_begin
    if x_init.ndim == 1:
        x_init = x_init[None, :]
    shared = x_init.shape[0] == 1 and P >= 1
    if not shared and x_init.shape[0] != P:
        raise ValueError("x_init must hold one initial state, or one per parameter set")
    arr = (_lib.sponet_cnvm_params * P)(*model_structs)
    cvs, cv_keep = cv_struct(cv_spec)
    D = 0 if cv_spec is None else cv_keep.dimension
    out_cv = _alloc((P, nrun, T, D), np.float64, x_init) if want_cv else None
    opts = _native_opts(seed, num_runs_total, run_begin, run_end, None, False, warps_per_chain, False, hist, chain_offset)
    sum_p = sumsq_p = None
    if sums is not None:
        sum_p, sumsq_p = _lib.ptr(sums[0]), _lib.ptr(sums[1])
    rc = L.sponet_cnvm_run_sweep(
        graph.handle if graph is not None else None, num_agents, arr, P, _lib.ptr(x_init), int(shared), _lib.ptr(t_eval), T,
        C.byref(opts), None if cvs is None else C.byref(cvs), None, _lib.ptr(out_cv), sum_p, sumsq_p, _lib.ptr(counters),
        _lib.current_stream_ptr(),
    )
    _lib.check(rc)
    return out_cv


# ---------------------------------------------------------------------------------------------
# replay mode
# ---------------------------------------------------------------------------------------------
def stream_budget(expected_events: float, draws_per_event: float) -> int:
    """uint64 values to pull for ``expected_events`` loop iterations (Poisson spread + ziggurat retries)."""
    e = max(expected_events, 1.0)
    return int(math.ceil(draws_per_event * 1.03 * (e + 10.0 * math.sqrt(e)))) + 4096


class RawStream:
    """Replays a numpy BitGenerator: pulls raw uint64s, and afterwards leaves the generator advanced
    by exactly the number of values the kernel consumed -- as if the reference loop had run on it."""

    def __init__(self, rng: Generator):
        self.bitgen = rng.bit_generator
        self.saved = self.bitgen.state

    def pull(self, n: int) -> np.ndarray:
        self.bitgen.state = self.saved
        return self.bitgen.random_raw(n)

    def commit(self, used: int) -> None:
        self.bitgen.state = self.saved
        if used:
            self.bitgen.random_raw(used)


def run_replay(
    kind: str,
    graph: GraphHandle | None,
    num_agents: int,
    model_struct,
    x_init: np.ndarray,
    num_runs: int,
    t_eval: np.ndarray,
    rngs: list[Generator],
    work: list[tuple[int, int, int, int]],
    per_run_budget: int,
    want_x: bool = True,
):
    """Bit-exact replay of W workers, worker w consuming ``rngs[w]`` over its (state, run) block.

    Returns (out_t [S,R,T], out_x [S,R,T,N] or None, events per worker).
    """
    L = _lib.lib()
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    S, T, W = x_init.shape[0], t_eval.size, len(work)
    streams = [RawStream(r) for r in rngs]
    out_t = np.empty((S, num_runs, T), dtype=np.float64)
    out_x = np.empty((S, num_runs, T, num_agents), dtype=x_init.dtype) if want_x else None
    used = np.zeros(W, dtype=np.int64)
    events = np.zeros(W, dtype=np.int64)
    scale = 1
    while True:
        lens = [max(1, (sb_e - sb) * (rb_e - rb)) * per_run_budget * scale for (sb, sb_e, rb, rb_e) in work]
        raw = np.concatenate([s.pull(n) for s, n in zip(streams, lens)])
        offs = np.concatenate([[0], np.cumsum(lens)])
        warr = (_lib.sponet_replay_work * W)(
            *[_lib.sponet_replay_work(sb, se, rb, re, int(offs[w]), int(lens[w])) for w, (sb, se, rb, re) in enumerate(work)]
        )
        args_tail = (
            _lib.ptr(x_init), S, num_runs, _lib.ptr(t_eval), T, _lib.ptr(raw), C.cast(warr, C.c_void_p), W,
            _lib.ptr(out_t), _lib.ptr(out_x), _lib.ptr(used), _lib.ptr(events), None,
        )
        if kind == "cnvm":
            rc = L.sponet_cnvm_replay(graph.handle if graph is not None else None, num_agents, C.byref(model_struct), *args_tail)
        else:
            rc = L.sponet_cntm_replay(graph.handle, C.byref(model_struct), *args_tail)
        try:
            _lib.check(rc)
            break
        except _lib.StreamExhausted:
            scale *= 2
            if scale > 1 << 12:
                raise
    for s, u in zip(streams, used):
        s.commit(int(u))
    return out_t, out_x, events


def run_replay_all(kind, graph, num_agents, model_struct, x_init: np.ndarray, t_max: float, rng: Generator,
                   expected_events: float, draws_per_event: float):
    """Store-every-change replay of one realization -> (t_traj, x_traj) like the reference lists."""
    L = _lib.lib()
    stream = RawStream(rng)
    budget = stream_budget(expected_events, draws_per_event)
    cap = int(expected_events + 10 * math.sqrt(max(expected_events, 1.0))) + 1024
    while True:
        raw = stream.pull(budget)
        log_t = np.empty(cap, dtype=np.float64)
        log_a = np.empty(cap, dtype=np.int32)
        log_o = np.empty(cap, dtype=np.int32)
        cnt, used, events = (np.zeros(1, dtype=np.int64) for _ in range(3))
        tail = (_lib.ptr(x_init), float(t_max), _lib.ptr(raw), raw.size, cap, _lib.ptr(log_t), _lib.ptr(log_a),
                _lib.ptr(log_o), _lib.ptr(cnt), _lib.ptr(used), _lib.ptr(events), None)
        if kind == "cnvm":
            rc = L.sponet_cnvm_replay_all(graph.handle if graph is not None else None, num_agents, C.byref(model_struct), *tail)
        else:
            rc = L.sponet_cntm_replay_all(graph.handle, C.byref(model_struct), *tail)
        try:
            _lib.check(rc)
            break
        except _lib.StreamExhausted:
            budget *= 2
        except ValueError as err:
            if "capacity" not in str(err):
                raise
            cap *= 2
    stream.commit(int(used[0]))
    k = int(cnt[0])
    t_traj = np.concatenate(([0.0], log_t[:k]))
    # the reference stores a full copy of x per accepted change (cnvm/model.py:232-234)
    x_traj = np.empty((k + 1, num_agents), dtype=x_init.dtype)
    x = x_init.copy()
    x_traj[0] = x
    for i in range(k):
        x[log_a[i]] = log_o[i]
        x_traj[i + 1] = x
    return t_traj, x_traj
