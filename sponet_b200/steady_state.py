"""Steady-state estimate of a collective variable (reference: sponet/steady_state.py:12-81).

The estimator is the reference's: burn in for 1000 characteristic times, then average the CV over
samples spaced one characteristic time apart until ``4 var / tol^2`` samples are in.

What changes is the shape of the work.  The reference walks ONE chain and evaluates the CV on the host after
every chunk.  Here ``num_chains`` independent chains (default 128) run side by side: one launch burns all of
them in (only the final states come back: ``final_state_only``), every further launch advances all of them by
up to ``max_chunk_size`` characteristic times while the simulation kernel evaluates the CV at every sample time
and accumulates sum c and sum c^2 with integer atomics -- no trajectory leaves the GPU.  Under
torch.distributed the chains are dealt to the ranks and the sums meet in one all-reduce per round.  The
stopping rule, the signature and the return shape ``(D,)`` are the reference's; ``rng`` is an extension (the
reference uses an unseedable module-level generator), ``verbose=True`` reproduces its progress prints.
Collective variables the kernels cannot accumulate exactly (float weights, Propensities, user classes),
replay mode and per-run networks take the reference's single-chain loop on ``model.simulate``.
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


def _multi_chain(model, kind, spec, delta_t, tol_abs, max_chunk_size, rng, say, num_chains):
    import torch

    from . import engine
    from .multiprocessing import _broadcast_seed, _dist, _split_runs
    from .sample_moments import _allreduce_device_

    p = model.params
    N, M, D = p.num_agents, p.num_opinions, spec.dimension
    rank, world = _dist()
    seed = engine.draw_seed(rng)
    if world > 1:
        seed = _broadcast_seed(seed)
    b = np.concatenate(([0], np.cumsum(_split_runs(num_chains, world))))
    lo, hi = int(b[rank]), int(b[rank + 1])
    dev = torch.device("cuda", torch.cuda.current_device())
    # one random initial state per chain, like simulate(x_init=None) (cnvm/model.py:100-103); derived from the
    # job-wide seed so that the result does not depend on the number of ranks
    states = np.random.default_rng(seed).integers(0, M, (num_chains, N), dtype=np.uint8)[lo:hi]
    x = torch.as_tensor(np.ascontiguousarray(states), device=dev)
    call = [0]

    def advance(x, t_eval, sums):
        """All local chains from ``x`` over ``t_eval``; chain ids are fresh in every call."""
        call[0] += 1
        kw = dict(cv_spec=spec if sums is not None else None, sums=sums, want_x=True, final_state_only=True,
                  chain_offset=call[0] * num_chains + lo)
        if kind == "cnvm":
            struct, keep = model._struct()
            out, _ = engine.run_native("cnvm", model._graph(), N, struct, x, t_eval, seed, 1, 0, 1, **kw)
            del keep
        else:
            out, _ = engine.run_native("cntm", model._graph(), N, model._struct(), x, t_eval, seed, 1, 0, 1, **kw)
        return out[:, 0]

    say("Integrating through transient phase...")
    if hi > lo:
        x = advance(x, np.array([0.0, 1000 * delta_t]), None)
    divisor = torch.as_tensor(spec.divisor_vector(), device=dev)
    tot = torch.zeros((2, D), dtype=torch.int64, device=dev)  # job-wide sum c, sum c^2 over all samples so far
    samples, remaining = 0, 1000
    while remaining > 0:
        per_chain = int(min(-(-remaining // num_chains), max_chunk_size))
        say(f"Acquiring {per_chain * num_chains} more samples...")
        say("")
        part = torch.zeros((2, D), dtype=torch.int64, device=dev)
        if hi > lo:
            sums = torch.zeros((2, hi - lo, per_chain + 1, D), dtype=torch.int64, device=dev)
            x = advance(x, delta_t * np.arange(per_chain + 1), (sums[0], sums[1]))
            part = sums[:, :, 1:].sum(dim=(1, 2))  # the sample at t = 0 of a round is the previous round's last one
        if world > 1:
            _allreduce_device_([part])
        tot += part
        tot_global = tot
        samples += per_chain * num_chains
        mean = tot_global[0].double() / samples / divisor
        var = tot_global[1].double() / samples / (divisor * divisor) - mean * mean
        estimate = mean.cpu().numpy()
        needed = int(4 * float(var.max().item()) / tol_abs**2)
        remaining = needed - samples
        if remaining > 0:
            say(f"Current esimate: {estimate}")
            say(f"Approximately {remaining} more samples are required.")
    say(f"Finished after acquiring {samples} samples.")
    say(f"Final estimate: {estimate}")
    return estimate


def estimate_steady_state(
    params: Parameters,
    col_var: CollectiveVariable,
    tol_abs: float = 0.01,
    max_chunk_size: int = 1000,
    rng: Generator | None = None,
    verbose: bool = True,
    num_chains: int | None = None,
) -> np.ndarray:
    """Time-average of ``col_var`` in the stationary regime, shape ``(col_var.dimension,)``.

    ``num_chains`` (extension): independent chains averaged together, default 128; ``num_chains=1`` with a CV the
    kernel cannot evaluate falls back to the reference's single-chain loop."""
    from . import engine
    from .multiprocessing import _kernel_cv_spec

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

    kind = "cnvm" if isinstance(params, CNVMParameters) else "cntm"
    spec = _kernel_cv_spec(col_var, params.num_agents) if params.num_opinions <= 256 else None
    if spec is not None and (spec.grouped or spec.divisors is not None) and kind != "cnvm":
        spec = None
    if spec is not None and (not spec.integer_valued or spec.divisors is not None or (spec.edge and params.num_opinions != 2)):
        spec = None
    if (spec is not None and engine.resolve_mode(None) == "native" and getattr(params, "network_generator", None) is None
            and (kind == "cntm" or not model.is_complete or params.num_agents <= 1 << 22)):
        return _multi_chain(model, kind, spec, delta_t, tol_abs, max_chunk_size, rng, say,
                            128 if num_chains is None else int(num_chains))

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
