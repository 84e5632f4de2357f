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
        sums = (np.zeros((1, T, spec.dimension), dtype=np.int64), np.zeros((1, T, spec.dimension), dtype=np.int64))
        x0 = np.ascontiguousarray(np.asarray(state, dtype=np.uint8)[None, :])
        if kind == "cnvm":
            struct, keep = model._struct()
            if plain and not spec.edge and use_count_only(model, kind, spec, batch_size):
                c0 = np.bincount(x0[0], minlength=p.num_opinions).astype(np.int64)[None, :]
                engine.run_native("cnvm_counts", None, N, struct, None, t_grid, seed, batch_size, rb, re,
                                  cv_spec=spec, sums=sums, counts_init=c0)
            else:
                engine.run_native("cnvm", model._graph(), N, struct, x0, t_grid, seed, batch_size, rb, re,
                                  cv_spec=spec, sums=sums)
            del keep
        else:
            engine.run_native("cntm", model._graph(), N, model._struct(), x0, t_grid, seed, batch_size, rb, re,
                              cv_spec=spec, sums=sums)
        if world > 1:
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
