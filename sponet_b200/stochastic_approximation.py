"""Well-mixed opinion-share sampler (reference: sponet/cnvm/approximations/stochastic_approximation.py:11-157),
routed through the count-only simulation kernel.

The reference runs Gillespie's direct method on the shares with propensities
``N c[m] (r[m,n] c[n] + r_tilde[m,n])`` -- the CNVM on a complete network in which an agent may also pick itself
as the neighbour it imitates.  Here the same ensemble is produced by ``sponet_cnvm_run_complete_counts``: the
exact CNVM on the complete graph with alpha = 1 (sponet/cnvm/model.py:365-433; state = the M opinion counts,
one GPU thread per realization), whose imitation propensity is ``N c[m] r[m,n] c[n] N / (N - 1)``: the two
processes differ by the factor N / (N - 1) on the imitation term, i.e. by O(1 / N) -- below the
approximation's own error on any network that is not complete, and statistically invisible for N >= 100
(tests/test_gpu_states.py compares both against each other).  Signature and return shapes are the reference's."""
from __future__ import annotations

import numpy as np
from numpy.random import Generator, default_rng
from numpy.typing import ArrayLike, NDArray

from . import engine
from .cnvm.model import CNVM
from .cnvm.parameters import CNVMParameters
from .utils import counts_from_shares, t_eval_to_ndarray


def sample_stochastic_approximation(
    params: CNVMParameters,
    initial_states: ArrayLike,
    t_max: float,
    num_samples: int,
    t_eval: ArrayLike,
    rng: Generator | None = None,
    seed: int | None = None,
) -> tuple[NDArray, NDArray]:
    """``initial_states``: opinion SHARES, shape (M,) or (S, M).  Returns ``(t, c)`` with c the shares,
    shape (S, num_samples, T, M), or (num_samples, T, M) for a single initial state."""
    if rng is not None:
        key = engine.draw_seed(rng)
    elif seed is not None:
        key = engine.draw_seed(default_rng(seed))
    else:
        key = engine.draw_seed(default_rng())
    t = t_eval_to_ndarray(t_eval, t_max)
    shares = np.array(initial_states, ndmin=1, dtype=float)
    is_1d = shares.ndim == 1
    if is_1d:
        shares = shares[None, :]
    M, N = params.num_opinions, params.num_agents
    if shares.shape[1] != M:
        raise ValueError(f"initial_states must have {M} shares per state")
    if M > 16:
        raise ValueError("the count-only kernel supports num_opinions <= 16")
    counts0 = np.ascontiguousarray(counts_from_shares(shares, N), dtype=np.int64)
    well_mixed = CNVMParameters(num_opinions=M, num_agents=N, r=params.r, r_tilde=params.r_tilde, alpha=1.0)
    model = CNVM(well_mixed)
    struct, keep = model._struct()
    spec = engine.KernelCV(np.arange(M, dtype=np.int32), float(N))
    _, c = engine.run_native("cnvm_counts", None, N, struct, None, t, key, num_samples, 0, num_samples, cv_spec=spec,
                             want_cv=True, counts_init=counts0)
    del keep
    return t, (c[0] if is_1d else c)
