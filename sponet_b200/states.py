"""Initial-state samplers on the GPU (reference: sponet/states.py:45-164).

Same names, arguments and return shapes as the reference -- ``(num_states, num_agents)`` or ``(num_agents,)`` for a
single state, opinions as integers -- but the states are drawn by ``sponet_sample_states`` on the device (Philox keyed
by one uint64 taken from ``rng``): agents uniform in {0..M-1}, or a uniformly random arrangement of given opinion
counts (draws without replacement; the reference shuffles ``np.repeat(opinions, counts)``).  The target counts
(``counts_from_shares``, Dirichlet shares) stay host arithmetic, as in the reference.

``device=True`` (extension) returns the states as a CUDA uint8 tensor ``(num_states, num_agents)``, which
``sample_many_runs`` accepts directly: a sweep over many initial states then has no O(S * N) host step at all.
The cluster / target-CV samplers of the reference (``sample_states_local_clusters``, ``sample_states_target_cvs``)
are graph searches on the host and out of scope.
"""
from __future__ import annotations

import numpy as np
from numpy.random import Generator, default_rng
from numpy.typing import ArrayLike

from . import _lib, engine
from .utils import counts_from_shares


def _draw(num_agents: int, num_opinions: int, counts, num_states: int, rng: Generator, device: bool):
    """One call of the C ABI: uint8 [num_states, num_agents] (numpy, or a CUDA tensor with ``device``)."""
    _lib.require_device()
    if num_opinions > 256:
        raise ValueError("the GPU state samplers hold opinions as uint8: num_opinions <= 256")
    seed = engine.draw_seed(rng)
    if device:
        import torch

        out = torch.empty((num_states, num_agents), dtype=torch.uint8, device="cuda")
    else:
        out = np.empty((num_states, num_agents), dtype=np.uint8)
    c = None
    shared = 0
    if counts is not None:
        c = np.ascontiguousarray(np.atleast_2d(counts), dtype=np.int64)
        shared = int(c.shape[0] == 1 and num_states >= 1)
    _lib.check(_lib.lib().sponet_sample_states(int(num_agents), int(num_opinions), _lib.ptr(c), shared, int(num_states),
                                               seed, _lib.ptr(out), _lib.current_stream_ptr()))
    return out


def _sample_states(func, num_states: int, unique: bool, device: bool):
    """sponet/states.py:15-42: squeeze a single state, redraw duplicates until ``num_states`` are unique."""
    states = func(num_states)
    if device:
        if unique and num_states > 1:  # duplicates are compared through a 64-bit hash of every state
            import torch

            w = torch.arange(1, states.shape[1] + 1, device=states.device, dtype=torch.int64) * 0x9E3779B97F4A7C15 % (1 << 61)
            h = (states.to(torch.int64) * w).sum(1)
            if torch.unique(h).numel() < num_states:
                host = _sample_states(lambda k: func(k).cpu().numpy(), num_states, True, False)
                return torch.as_tensor(np.ascontiguousarray(host, dtype=np.uint8), device=states.device)
        return states
    if states.shape[0] == 1:
        return states[0]
    if not unique:
        return states
    states = np.unique(states, axis=0)
    while states.shape[0] < num_states:
        missing = num_states - states.shape[0]
        states = np.unique(np.concatenate([states, func(missing)]), axis=0)
        if num_states - states.shape[0] == missing:
            from warnings import warn

            warn("Sampling of unique states might be in an endless loop.")
    return states


def sample_states_uniform(num_agents: int, num_opinions: int, num_states: int = 1, rng: Generator = default_rng(),
                          unique: bool = True, device: bool = False):
    """Every agent's opinion uniform in {0, ..., num_opinions - 1} (sponet/states.py:45-77)."""
    return _sample_states(lambda k: _as_int(_draw(num_agents, num_opinions, None, k, rng, device), device), num_states, unique, device)


def sample_states_uniform_shares(num_agents: int, num_opinions: int, num_states: int = 1, rng: Generator = default_rng(),
                                 unique: bool = True, device: bool = False):
    """Opinion shares uniform on the simplex (Dirichlet(1, ..., 1)), each state randomly arranged
    (sponet/states.py:80-121)."""
    def sample(k):
        shares = rng.dirichlet(np.ones(num_opinions), k)
        counts = counts_from_shares(shares, num_agents)
        return _as_int(_draw(num_agents, num_opinions, counts, k, rng, device), device)

    return _sample_states(sample, num_states, unique, device)


def sample_states_target_shares(num_agents: int, target_shares: ArrayLike, num_states: int = 1,
                                rng: Generator = default_rng(), unique: bool = True, device: bool = False):
    """Every state has exactly ``counts_from_shares(target_shares, num_agents)`` agents per opinion, randomly arranged
    (sponet/states.py:124-164)."""
    counts = counts_from_shares(target_shares, num_agents)
    return _sample_states(lambda k: _as_int(_draw(num_agents, counts.shape[0], counts[None, :], k, rng, device), device),
                          num_states, unique, device)


def _as_int(x, device: bool):
    """The reference returns int64 arrays (rng.integers / np.repeat of an arange); device tensors stay uint8."""
    return x if device else x.astype(np.int64)
