"""The one sampling primitive of the reference with arithmetic worth sharing (reference: sponet/sampling.py:54-96).

The alias-table build runs in the C++ part of libsponet_b200 with the reference's sequential float64 sums and stack
order, and is the table the replay kernels consume.  The per-draw helpers of the reference (`sample_randint`,
`sample_from_alias`, ...) are njit one-liners inlined into its loops; here they live inside the kernels
(`replay.cu`, `native_cnvm.cu`) and have no host-side twin.
"""
from __future__ import annotations

import numpy as np
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
