"""Batched moment estimation (reference: sponet/sample_moments.py:16-128).

Same signature, stopping rule (``4 sigma / sqrt(n) / mean < rel_tol``), ``.npz`` keys and return tuple.
The batch loop is the reference's; what differs is where the sums come from:

* unweighted ``OpinionShares`` (the common case): the simulation kernel accumulates ``sum c`` and
  ``sum c^2`` over the runs with integer atomics and never materialises the per-run trajectories;
* any other CV / no CV: the per-run array from ``sample_many_runs`` is reduced by ``sponet_moments``.

Under ``torch.distributed`` each rank simulates its share of every batch and the ``[2, T, D]`` partial
sums are all-reduced (NCCL over NVLink on GPUs, gloo on CPU test rigs) -- the only collective on the
whole path.
"""
from __future__ import annotations

import sys
import time
from typing import Optional

import numpy as np
from numpy.random import Generator, default_rng
from numpy.typing import ArrayLike, NDArray

from . import _lib, engine
from .collective_variables import CollectiveVariable
from .multiprocessing import _dist, _kernel_cv_spec, _make_model, _split_runs, sample_many_runs, use_count_only
from .parameters import Parameters


def _allreduce_device_(tensors: list) -> None:
    """In-place sum over ranks of device tensors (the integer partial sums the kernels accumulated on this GPU),
    packed into ONE message: NCCL all-reduce on device buffers -- no host bounce.  gloo test rigs pass CPU tensors."""
    import torch
    import torch.distributed as dist

    flat = torch.cat([t.reshape(-1).view(torch.int64) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    o = 0
    for t in tensors:
        t.view(torch.int64).reshape(-1).copy_(flat[o : o + t.numel()])
        o += t.numel()


def sweep_pieces(num_points: int, num_runs: int, rank: int, world: int, weights=None):
    """Shard of a P x R sweep for one rank: a contiguous range of the flattened (point, run) index, as at most
    three pieces [(point_begin, point_end, run_begin, run_end)] -- a partial first point, whole points, a partial
    last point.  Without ``weights`` the range is cut like `_split_runs` (equal numbers of chains); with
    ``weights`` (cost of one run of every point, e.g. its event rate) the cuts fall at equal shares of the
    total COST, so that ranks whose points are cheap take more of them -- the all-reduce then does not wait for
    the rank that drew the expensive corner of the grid.  P >= world gives (nearly) whole points per rank,
    P < world splits the runs of a point."""
    total = num_points * num_runs
    if weights is None:
        b = np.concatenate(([0], np.cumsum(_split_runs(total, world))))
    else:
        w = np.asarray(weights, dtype=np.float64)
        cum = np.concatenate(([0.0], np.cumsum(w * num_runs)))  # cost up to the start of every point
        b = [0]
        for k in range(1, world):
            target = cum[-1] * k / world
            pt = int(np.searchsorted(cum, target, side="right") - 1)
            pt = min(max(pt, 0), num_points - 1)
            run = int(round((target - cum[pt]) / w[pt])) if w[pt] > 0 else 0
            b.append(max(b[-1], min(total, pt * num_runs + min(max(run, 0), num_runs))))
        b.append(total)
        b = np.asarray(b)
    lo, hi = int(b[rank]), int(b[rank + 1])
    pieces = []
    while lo < hi:
        p, r = divmod(lo, num_runs)
        if r == 0 and hi - lo >= num_runs:  # whole points
            q = p + (hi - lo) // num_runs
            pieces.append((p, q, 0, num_runs))
            lo = q * num_runs
        else:
            r_end = min(num_runs, r + hi - lo)
            pieces.append((p, p + 1, r, r_end))
            lo += r_end - r
    return pieces


def sample_moments_sweep(
    params_list,
    initial_state: ArrayLike,
    t_max: float,
    num_timesteps: int,
    num_runs: int,
    collective_variable: CollectiveVariable,
    rng: Generator = default_rng(),
    hist_bins: int = 0,
    stats: dict | None = None,
):
    """Moments of a collective variable for a whole PARAMETER SWEEP in one pass (BASELINE config 5).

    The reference has no sweep entry point: a user loops `sample_moments` (sponet/sample_moments.py:16-128) over
    the points, one process pool per call.  Here the P rate sets (``params_list``: CNVMParameters on ONE network,
    same num_opinions and alpha) run as ONE kernel launch per GPU -- every chain looks its acceptance table up
    by its point -- the kernel accumulates sum c, sum c^2 (and, with ``hist_bins`` > 0, a histogram of c) per
    point with integer atomics, and under torch.distributed the points are dealt to the ranks and the integer
    partials meet in ONE all-reduce on device buffers.  ``num_runs`` realizations per point, no stopping rule.

    Returns ``(t, mean, variance)`` with shapes (T,), (P, T, D), (P, T, D); with ``hist_bins`` additionally
    ``(hist, edges)``: hist uint64 (P, T, D, bins) and the bins + 1 edges in units of the CV.
    Point p is bit-identical to a stand-alone run of its parameters with the same seed and chain ids."""
    import torch

    from .cnvm.model import CNVM
    from .cnvm.parameters import CNVMParameters
    from .multiprocessing import _broadcast_seed, _check_states

    params_list = list(params_list)
    P = len(params_list)
    if P == 0 or not all(isinstance(q, CNVMParameters) for q in params_list):
        raise ValueError("Parameters not valid.")
    p0 = params_list[0]
    for q in params_list:
        if q.num_opinions != p0.num_opinions or q.network is not p0.network or q.alpha != p0.alpha or q.num_agents != p0.num_agents:
            raise ValueError("a sweep runs on one network: all parameter sets must share network, num_opinions and alpha")
        if getattr(q, "network_generator", None) is not None:
            raise ValueError("a sweep needs a fixed network")
    model, _ = _make_model(p0)
    N = p0.num_agents
    spec = _kernel_cv_spec(collective_variable, N)
    if spec is None or not spec.integer_valued or (spec.edge and p0.num_opinions != 2):
        raise ValueError("sample_moments_sweep needs a collective variable the kernel accumulates exactly "
                         "(OpinionShares family with integer weights, Interfaces)")
    x0 = np.array(initial_state, ndmin=1)
    _check_states(x0, p0.num_opinions)
    x0 = np.ascontiguousarray(x0.astype(np.uint8))[None, :]
    t = np.linspace(0, t_max, num_timesteps)
    T, D = t.size, spec.dimension
    rank, world = _dist()
    seed = engine.draw_seed(rng)
    if world > 1:
        seed = _broadcast_seed(seed)
    dev = torch.device("cuda", torch.cuda.current_device())
    sums = torch.zeros((2, P, T, D), dtype=torch.int64, device=dev)
    hist = None
    if hist_bins:
        vmax = N if not spec.grouped else int(np.abs(spec.agent_weight).sum()) if spec.agent_weight is not None else N
        lo, width = engine.Histogram.for_counts(hist_bins, vmax)
        hist_t = torch.zeros((P, T, D, hist_bins), dtype=torch.int64, device=dev)
    da = model.degree_alpha
    keep, structs = [], []
    for q in params_list:
        st, k = engine.cnvm_struct(q, da)
        structs.append(st)
        keep.append(k)
    # every struct must point at ONE degree_alpha array (the C ABI checks pointer equality)
    da_arr = keep[0][2]
    for st in structs:
        st.degree_alpha = _lib.ptr(da_arr)
    d_x0 = torch.as_tensor(x0, device=dev)
    graph = model._graph()
    t0 = time.perf_counter()
    # cost of one run of every point: its expected events plus a fixed part per chain (copy of the packed initial
    # state, T snapshots; ~0.03 N event-equivalents, measured on the 256-point sweep of config 5)
    weights = [((q.r_imit * float(np.sum(da)) + q.r_noise * N) if p0.network is not None
                else N * (q.r_imit * float(da[0]) + q.r_noise)) * float(t_max) + 0.03 * N for q in params_list]
    for (pb, pe, rb, re) in sweep_pieces(P, num_runs, rank, world, weights):
        h = None if not hist_bins else engine.Histogram(hist_bins, lo, width, hist_t[pb:pe])
        engine.run_native_sweep(graph, N, structs[pb:pe], d_x0, t, seed, num_runs, rb, re, cv_spec=spec,
                                sums=(sums[0, pb:pe], sums[1, pb:pe]), hist=h, chain_offset=pb * num_runs)
    if stats is not None:
        torch.cuda.synchronize()
        stats["kernel_s"] = time.perf_counter() - t0
    if world > 1:
        t1 = time.perf_counter()
        _allreduce_device_([sums] + ([hist_t] if hist_bins else []))
        if stats is not None:
            torch.cuda.synchronize()
            stats["allreduce_s"] = time.perf_counter() - t1
            stats["allreduce_bytes"] = int(sums.numel() * 8 + (hist_t.numel() * 8 if hist_bins else 0))
    host = sums.cpu().numpy()
    del keep
    divisor = spec.divisor_vector()
    mean = host[0] / num_runs / divisor
    variance = host[1] / num_runs / (divisor * divisor) - mean**2
    if hist_bins:
        return t, mean, variance, hist_t.cpu().numpy().astype(np.uint64), engine.Histogram(hist_bins, lo, width, None).edges(spec.divisor)
    return t, mean, variance


def sample_histogram(
    params: Parameters,
    initial_state: ArrayLike,
    t_max: float,
    num_timesteps: int,
    num_runs: int,
    collective_variable: CollectiveVariable,
    bins: int = 64,
    rng: Generator = default_rng(),
):
    """Distribution of a collective variable over ``num_runs`` realizations at every output time, accumulated inside
    the simulation kernel (integer atomics per (time, component, bin); all-reduced over the ranks of a
    torch.distributed job together with the moment sums).  Returns ``(t, hist, edges, mean, variance)``:
    hist uint64 (T, D, bins) with hist.sum(-1) == num_runs, edges (bins + 1,) in units of the CV (the raw range
    0 .. N split into equal integer-width bins)."""
    import torch

    from .multiprocessing import _broadcast_seed, _check_states

    model, kind = _make_model(params)
    p = model.params
    N = p.num_agents
    spec = _kernel_cv_spec(collective_variable, N)
    if spec is not None and (spec.grouped or spec.divisors is not None) and kind != "cnvm":
        spec = None
    if spec is None or not spec.integer_valued or (spec.edge and p.num_opinions != 2) or spec.divisors is not None:
        raise ValueError("sample_histogram needs a collective variable the kernel accumulates exactly")
    if getattr(p, "network_generator", None) is not None:
        raise ValueError("sample_histogram needs a fixed network")
    x0 = np.array(initial_state, ndmin=1)
    _check_states(x0, p.num_opinions)
    x0 = np.ascontiguousarray(x0.astype(np.uint8))[None, :]
    t = np.linspace(0, t_max, num_timesteps)
    T, D = t.size, spec.dimension
    rank, world = _dist()
    seed = engine.draw_seed(rng)
    if world > 1:
        seed = _broadcast_seed(seed)
    b = np.concatenate(([0], np.cumsum(_split_runs(num_runs, world))))
    rb, re = int(b[rank]), int(b[rank + 1])
    dev = torch.device("cuda", torch.cuda.current_device())
    sums = torch.zeros((2, 1, T, D), dtype=torch.int64, device=dev)
    hist_t = torch.zeros((1, T, D, bins), dtype=torch.int64, device=dev)
    if spec.edge:
        vmax = int(spec.graph.col.size // 2)  # Interfaces: at most the number of edges
    elif spec.agent_weight is not None:
        vmax = int(np.abs(spec.agent_weight).sum())
    else:
        vmax = N
    lo, width = engine.Histogram.for_counts(bins, vmax)
    h = engine.Histogram(bins, lo, width, hist_t)
    d_x0 = torch.as_tensor(x0, device=dev)
    if kind == "cnvm":
        struct, keep = model._struct()
        engine.run_native("cnvm", model._graph(), N, struct, d_x0, t, seed, num_runs, rb, re, cv_spec=spec,
                          sums=(sums[0], sums[1]), hist=h)
        del keep
    else:
        engine.run_native("cntm", model._graph(), N, model._struct(), d_x0, t, seed, num_runs, rb, re, cv_spec=spec,
                          sums=(sums[0], sums[1]), hist=h)
    if world > 1:
        _allreduce_device_([sums, hist_t])
    host = sums.cpu().numpy()
    divisor = spec.divisor_vector()
    mean = host[0, 0] / num_runs / divisor
    variance = host[1, 0] / num_runs / (divisor * divisor) - mean**2
    return t, hist_t[0].cpu().numpy().astype(np.uint64), h.edges(spec.divisor), mean, variance


def _allreduce_(arrays: list[np.ndarray]) -> None:
    """In-place sum over ranks of equally-typed host arrays, packed into one message."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    flat = torch.from_numpy(np.concatenate([a.ravel() for a in arrays])).to(dev)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat = flat.cpu().numpy()
    o = 0
    for a in arrays:
        a[...] = flat[o : o + a.size].reshape(a.shape)
        o += a.size


def batch_moment_sums(model, kind, state, t_grid, batch_size, cv, rng) -> tuple[NDArray, NDArray]:
    """(sum_x, sum_xx) of shape (T, D) over one batch of ``batch_size`` runs (global, all ranks)."""
    p = model.params
    N, T = p.num_agents, t_grid.size
    rank, world = _dist()
    spec = _kernel_cv_spec(cv, N) if p.num_opinions <= 256 else None
    if spec is not None and (spec.grouped or spec.divisors is not None) and kind != "cnvm":
        spec = None
    if spec is not None and (not spec.integer_valued or (spec.edge and p.num_opinions != 2)):
        spec = None  # Propensities are float sums: moments from the returned CV values instead
    # A NetworkGenerator means a fresh network for every run (cnvm/model.py:95-96 via sample_many_runs): that is
    # what the generic path below does; the fused path would simulate one (stale or complete) network.
    per_run_network = getattr(p, "network_generator", None) is not None
    if spec is not None and not per_run_network and engine.resolve_mode(None) == "native":
        from .multiprocessing import _broadcast_seed

        seed = engine.draw_seed(rng)
        if world > 1:
            seed = _broadcast_seed(seed)
        b = np.concatenate(([0], np.cumsum(_split_runs(batch_size, world))))
        rb, re = int(b[rank]), int(b[rank + 1])
        divisor = spec.divisor_vector()
        plain = not spec.grouped and spec.divisors is None
        # Under NCCL the kernel accumulates into device tensors and those go into the all-reduce as they are (no host
        # bounce); a single process hands host arrays to the C ABI, which stages them.
        from .multiprocessing import _nccl

        on_device = world > 1 and _nccl()
        if on_device:
            import torch

            dsum = torch.zeros((2, 1, T, spec.dimension), dtype=torch.int64, device="cuda")
            sums = (dsum[0], dsum[1])
            x0 = torch.as_tensor(np.ascontiguousarray(np.asarray(state, dtype=np.uint8)[None, :]), device="cuda")
            x0_host = np.asarray(state, dtype=np.uint8)
        else:
            sums = (np.zeros((1, T, spec.dimension), dtype=np.int64), np.zeros((1, T, spec.dimension), dtype=np.int64))
            x0 = np.ascontiguousarray(np.asarray(state, dtype=np.uint8)[None, :])
            x0_host = x0[0]
        if kind == "cnvm":
            struct, keep = model._struct()
            if plain and not spec.edge and use_count_only(model, kind, spec, batch_size):
                c0 = np.bincount(x0_host, minlength=p.num_opinions).astype(np.int64)[None, :]
                if on_device:
                    c0 = torch.as_tensor(c0, device="cuda")
                engine.run_native("cnvm_counts", None, N, struct, None, t_grid, seed, batch_size, rb, re,
                                  cv_spec=spec, sums=sums, counts_init=c0)
            else:
                engine.run_native("cnvm", model._graph(), N, struct, x0, t_grid, seed, batch_size, rb, re,
                                  cv_spec=spec, sums=sums)
            del keep
        else:
            engine.run_native("cntm", model._graph(), N, model._struct(), x0, t_grid, seed, batch_size, rb, re,
                              cv_spec=spec, sums=sums)
        if on_device:
            _allreduce_device_([dsum])
            host = dsum.cpu().numpy()
            sums = (host[0], host[1])
        elif world > 1:
            _allreduce_([sums[0], sums[1]])
        # integer sums of counts -> sums of the (possibly normalised) CV
        return sums[0][0] / divisor, sums[1][0] / (divisor * divisor)

    # generic path: per-run values, then the moment kernel (rows added in run order, like np.sum(x, 0))
    _, x = sample_many_runs(p, np.array(state, ndmin=1), float(t_grid[-1]), t_grid if t_grid.size > 1 else 1,
                            batch_size, collective_variable=cv, rng=rng)
    x = np.ascontiguousarray(x, dtype=np.float64)
    L = x.shape[1] * x.shape[2]
    s, q = np.empty(L), np.empty(L)
    _lib.check(_lib.lib().sponet_moments(_lib.ptr(x), x.shape[0], L, _lib.ptr(s), _lib.ptr(q), None))
    return s.reshape(x.shape[1:]), q.reshape(x.shape[1:])


def sample_moments(
    params: Parameters,
    initial_state: ArrayLike,
    t_max: float,
    num_timesteps: int,
    batch_size: int,
    rel_tol: float = 0.01,
    n_jobs: Optional[int] = None,
    collective_variable: Optional[CollectiveVariable] = None,
    rng: Generator = default_rng(),
    filename: Optional[str] = None,
    timeout_seconds: Optional[int] = None,
    verbose: bool = False,
) -> tuple[NDArray, NDArray, NDArray, int]:
    """Sample batches until the 4-sigma interval of the mean is below ``rel_tol`` of the mean.

    Returns ``(t, mean, variance, num_samples)`` with ``mean.shape == (num_timesteps, D)``.
    ``n_jobs`` is accepted for compatibility and ignored (the batch runs on the GPU).
    """
    model, kind = _make_model(params)
    t = np.linspace(0, t_max, num_timesteps)
    dim = params.num_agents if collective_variable is None else collective_variable.dimension
    if timeout_seconds is None:
        timeout_seconds = sys.maxsize

    sum_x = np.zeros((num_timesteps, dim))
    sum_xx = np.zeros((num_timesteps, dim))
    num_samples = 0
    start = time.time()
    while True:
        if verbose:
            print(f"Total number of samples: {num_samples}.")
        s, q = batch_moment_sums(model, kind, initial_state, t, batch_size, collective_variable, rng)
        num_samples += batch_size
        sum_x += s
        sum_xx += q

        mean = sum_x / num_samples
        variance = sum_xx / num_samples - mean**2
        with np.errstate(invalid="ignore", divide="ignore"):
            confidence_size = 4 * np.sqrt(variance / num_samples)
            rel_error = np.nanmax(confidence_size / mean)
        if filename is not None:
            np.savez_compressed(
                filename, t=t, mean=mean, variance=variance, num_samples=num_samples,
                confidence_size=confidence_size, rel_error=rel_error, time_elapsed=time.time() - start,
            )
        if verbose:
            print(f"Relative error: {rel_error}.\n")
        if rel_error < rel_tol or time.time() - start >= timeout_seconds:
            return t, mean, variance, num_samples
