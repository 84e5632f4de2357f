"""Sampling primitives exposed to users (reference: sponet/sampling.py).

The alias-table build is the only one with arithmetic worth sharing: it runs in the C++ part of
libsponet_b200 with the reference's sequential float64 sums and stack order, and is the table the
replay kernels consume.
"""
from __future__ import annotations

import numpy as np
from numpy.random import Generator
from numpy.typing import NDArray

from . import _lib


def build_alias_table(weights: NDArray) -> tuple[NDArray, NDArray]:
    """Vose alias tables (prob float64[n], alias int32[n]) for ``weights`` (sponet/sampling.py:54-96)."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.ndim != 1 or w.size == 0:
        raise ValueError("weights must be a non-empty 1-D array")
    prob = np.empty_like(w)
    alias = np.empty(w.size, dtype=np.int32)
    _lib.check(_lib.lib().sponet_build_alias_table(_lib.ptr(w), w.size, _lib.ptr(prob), _lib.ptr(alias)))
    return prob, alias


def sample_randint(high_excl: int, rng: Generator) -> int:
    """``int(rng.random() * high_excl)`` (sponet/sampling.py:7-25)."""
    return int(rng.random() * high_excl)


def sample_randint_other(high_excl: int, this: int, rng: Generator) -> int:
    """Uniform integer in ``[0, high_excl)`` without ``this`` (sponet/sampling.py:28-51)."""
    i = int(rng.random() * (high_excl - 1))
    return i if i < this else i + 1


def sample_from_alias(table_prob: NDArray, table_alias: NDArray, rng: Generator) -> int:
    """One-uniform alias sample (sponet/sampling.py:99-122)."""
    u = rng.random()
    n = table_prob.shape[0]
    idx = int(u * n)
    return idx if n * u - idx < table_prob[idx] else int(table_alias[idx])


def sample_weighted_bisect(prob_cum_sum: NDArray, rng: Generator) -> int:
    """Inverse-CDF sample by bisection (sponet/sampling.py:125-146)."""
    return int(np.searchsorted(prob_cum_sum, rng.random(), side="right"))
