"""Steady-state estimate of a collective variable (reference: sponet/steady_state.py:12-81).

The estimator is the reference's: burn in for 1000 characteristic times, then average the CV over
samples spaced one characteristic time apart until ``4 var / tol^2`` samples are in.  The chain runs
on the GPU through ``model.simulate``; ``rng`` is an extension (the reference uses an unseedable
module-level generator) and ``verbose=True`` reproduces its progress prints.
"""
from __future__ import annotations

import numpy as np
from numpy.random import Generator, default_rng

from .cntm.model import CNTM
from .cntm.parameters import CNTMParameters
from .cnvm.model import CNVM
from .cnvm.parameters import CNVMParameters
from .collective_variables import CollectiveVariable
from .parameters import Parameters


def estimate_steady_state(
    params: Parameters,
    col_var: CollectiveVariable,
    tol_abs: float = 0.01,
    max_chunk_size: int = 1000,
    rng: Generator | None = None,
    verbose: bool = True,
) -> np.ndarray:
    """Time-average of ``col_var`` in the stationary regime, shape ``(col_var.dimension,)``."""
    if isinstance(params, CNVMParameters):
        delta_t = 1.0 / (np.max(params.r) + params.num_opinions * np.max(params.r_tilde))
        model = CNVM(params)
    elif isinstance(params, CNTMParameters):
        delta_t = 1.0 / (params.r + params.r_tilde)
        model = CNTM(params)
    else:
        raise ValueError("Parameters not valid.")
    rng = default_rng() if rng is None else rng
    say = print if verbose else (lambda *a, **k: None)
    delta_t = float(delta_t)

    say("Integrating through transient phase...")
    _, x = model.simulate(1000 * delta_t, t_eval=2, rng=rng)

    samples = col_var(x)[1:]
    estimate = samples[-1]
    remaining = 1000
    while remaining > 0:
        remaining = min(remaining, max_chunk_size)
        say(f"Acquiring {remaining} more samples...")
        say("")
        _, x = model.simulate(remaining * delta_t, x[-1], t_eval=remaining + 1, rng=rng)
        samples = np.concatenate([samples, col_var(x)[1:]], axis=0)
        estimate = np.mean(samples, axis=0)
        needed = int(4 * np.max(np.var(samples, axis=0)) / tol_abs**2)
        remaining = needed - samples.shape[0]
        if remaining > 0:
            say(f"Current esimate: {estimate}")
            say(f"Approximately {remaining} more samples are required.")
    say(f"Finished after acquiring {samples.shape[0]} samples.")
    say(f"Final estimate: {estimate}")
    return estimate
