"""Host-side helpers of the ensemble path (time grids, nearest-index matching, adjacency -> CSR).

Mirrors the public helpers of the reference's ``sponet/utils.py`` that sit on the hot path's
boundary; each function cites the behaviour it reproduces.
"""
from __future__ import annotations

import numpy as np
from numpy.typing import ArrayLike, NDArray


def t_eval_to_ndarray(t_eval: ArrayLike, t_max: float) -> NDArray:
    """Normalise ``t_eval`` to a float64 grid (reference: sponet/utils.py:173-190).

    ``int n`` -> ``linspace(0, t_max, n)``; a float is rejected; arrays must be non-empty, strictly
    increasing, >= 0 and <= t_max.  Same ValueError messages as the reference.
    """
    if isinstance(t_eval, float):
        raise ValueError("t_eval has to be an array of time points or an int.")
    if isinstance(t_eval, int):
        return np.linspace(0, t_max, t_eval)
    grid = np.array(t_eval, dtype=float)
    if len(grid) == 0:
        raise ValueError("t_eval cannot be empty.")
    if np.min(np.diff(grid)) <= 0:
        raise ValueError("The times in t_eval have to be increasing.")
    if grid[0] < 0:
        raise ValueError("The times in t_eval have to be >= 0.")
    if grid[-1] > t_max:
        raise ValueError("The times in t_eval cannot be larger than t_max.")
    return grid


def argmatch(x_ref: NDArray, x: NDArray) -> NDArray:
    """Indices ``i`` minimising ``|x[i] - x_ref[k]|`` for sorted 1-D inputs (sponet/utils.py:7-45).

    Ties go to the larger index, values of ``x_ref`` beyond ``x[-1]`` map to the last index, exactly
    like the reference's merge scan.
    """
    x_ref = np.asarray(x_ref, dtype=float)
    x = np.asarray(x, dtype=float)
    out = np.zeros(x_ref.shape[0], dtype=np.int64)
    k, i, last = 0, 0, x.shape[0] - 1
    while x_ref[k] < x[i]:
        k += 1
    while k < x_ref.shape[0] and i < last:
        lo, hi, ref = x[i], x[i + 1], x_ref[k]
        if lo <= ref <= hi:
            out[k] = i if abs(lo - ref) < abs(hi - ref) else i + 1
            k += 1
        else:
            i += 1
    out[k:] = last
    return out


def mask_subsequent_duplicates(x: NDArray) -> NDArray:
    """Boolean mask dropping entries equal to their predecessor along axis 0 (sponet/utils.py:48-73)."""
    x = np.asarray(x)
    if x.ndim == 1:
        changed = x[1:] != x[:-1]
    elif x.ndim == 2:
        changed = (x[1:] != x[:-1]).any(axis=1)
    else:
        raise ValueError("Only 1D and 2D arrays are supported.")
    return np.concatenate(([True], changed))


def calculate_neighbor_list(network) -> list[NDArray]:
    """Adjacency list in node order, int32, neighbour order as networkx yields it (sponet/utils.py:76-93)."""
    return [np.fromiter(network.neighbors(v), dtype=np.int32) for v in network.nodes()]


def neighbor_list_to_csr(neighbor_list) -> tuple[NDArray, NDArray]:
    """CSR (row_ptr int32[n+1], col int32[nnz]) with the row order preserved.

    The k-th entry of a row is what ``int(u * deg)`` selects in the reference loop
    (sponet/cnvm/model.py:285-286), so the order is part of the replay contract.
    """
    n = len(neighbor_list)
    deg = np.fromiter((len(nb) for nb in neighbor_list), dtype=np.int64, count=n)
    if deg.sum() >= 2**31:
        raise ValueError("network too large for int32 CSR offsets")
    row_ptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(deg, out=row_ptr[1:])
    if row_ptr[-1] == 0:
        return row_ptr, np.zeros(0, dtype=np.int32)
    col = np.concatenate([np.asarray(nb, dtype=np.int32).ravel() for nb in neighbor_list])
    return row_ptr, np.ascontiguousarray(col, dtype=np.int32)


def counts_from_shares(shares: ArrayLike, num_agents: int) -> NDArray:
    """Integer counts summing to ``num_agents`` closest to ``shares * num_agents`` (sponet/utils.py:138-170).

    Floors every entry, then hands the missing agents to the entries with the largest remainders
    (``np.argpartition`` on the negative remainders, as the reference does).
    """
    scaled = np.array(shares) * num_agents
    if scaled.ndim == 1:
        return _round_preserving_sum(scaled, num_agents)
    out = np.zeros_like(scaled, dtype=int)
    for i, row in enumerate(scaled):
        out[i] = _round_preserving_sum(row, num_agents)
    return out


def _round_preserving_sum(scaled: NDArray, total: int) -> NDArray:
    base = np.floor(scaled)
    deficit = base - scaled
    base = base.astype(int)
    missing = total - int(base.sum())
    bump = np.argpartition(deficit, missing)[:missing]
    base[bump] += 1
    return base
