"""Collective variables (reference: sponet/collective_variables.py).

The OpinionShares family is a (weighted) histogram of opinions per snapshot.  Calling a CV on an
array of states runs the histogram kernel of libsponet_b200 (``sponet_opinion_shares``); inside
``sample_many_runs`` the unweighted OpinionShares is not called at all -- the simulation kernel keeps
the counts incrementally and emits them at the output times.

Every class keeps the reference's constructor, ``dimension`` attribute and ``__call__`` contract:
``(num_agents,) -> (dimension,)`` and ``(num_states, num_agents) -> (num_states, dimension)``.
"""
from __future__ import annotations

from typing import Protocol

import numpy as np
from numpy.typing import ArrayLike, NDArray

from . import _lib


class CollectiveVariable(Protocol):
    dimension: int

    def __call__(self, x: NDArray) -> NDArray: ...


def _weighted_histogram(x, num_opinions: int, weights: NDArray | None):
    """``out[s, m] = sum_j weights[j] * [x[s, j] == m]`` on the GPU.

    ``x`` is a host array (S, N) or a CUDA ``torch.Tensor``; the result lives where ``x`` lives.
    Reference: ``_opinion_shares_numba`` (sponet/collective_variables.py:400-405).
    """
    if num_opinions > 256:
        raise NotImplementedError("histogram kernel supports num_opinions <= 256")
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    if isinstance(x, np.ndarray):
        if x.size and (x.min() < 0 or x.max() >= num_opinions):
            # np.bincount would grow the output / raise; the reference then fails on assignment
            raise ValueError("state contains opinions outside [0, num_opinions)")
        xs = np.ascontiguousarray(x, dtype=np.uint8)
        out = np.empty((xs.shape[0], num_opinions), dtype=np.float64)
        _lib.check(
            _lib.lib().sponet_opinion_shares(
                _lib.ptr(xs), xs.shape[0], xs.shape[1], num_opinions, _lib.ptr(w), _lib.ptr(out), None
            )
        )
        return out
    import torch  # device tensor path

    xs = x.contiguous()
    if xs.dtype != torch.uint8:
        xs = xs.to(torch.uint8)
    out = torch.empty((xs.shape[0], num_opinions), dtype=torch.float64, device=xs.device)
    with torch.cuda.device(xs.device):
        _lib.check(
            _lib.lib().sponet_opinion_shares(
                _lib.ptr(xs), xs.shape[0], xs.shape[1], num_opinions, _lib.ptr(w), _lib.ptr(out),
                _lib.current_stream_ptr(),
            )
        )
    return out


def _as_2d(x):
    """(states_2d, was_1d) -- the reference's ``handle_1d`` decorator (collective_variables.py:40-60)."""
    if isinstance(x, np.ndarray) or not hasattr(x, "data_ptr"):
        x = np.asarray(x)
    if x.ndim == 1:
        return x[None, :], True
    return x, False


def _select(idx_to_return, num_opinions):
    if idx_to_return is None:
        return np.arange(num_opinions)
    return np.atleast_1d(np.array(idx_to_return))


class OpinionShares:
    """Counts (or shares) of each opinion, optionally weighted per agent.

    Reference: sponet/collective_variables.py:63-124.  ``normalize`` divides by ``num_agents`` when
    unweighted and by ``sum(|weights|)`` otherwise; ``idx_to_return`` selects opinions.
    """

    def __init__(
        self,
        num_opinions: int,
        normalize: bool = False,
        weights: ArrayLike | None = None,
        idx_to_return: ArrayLike | None = None,
    ):
        self.num_opinions = num_opinions
        self.weights = None if weights is None else np.array(weights)
        self.normalize = normalize
        self.normalization = None if weights is None else np.sum(np.abs(weights))
        self.idx_to_return = _select(idx_to_return, num_opinions)
        self.dimension = len(self.idx_to_return)

    def __call__(self, x):
        xs, was_1d = _as_2d(x)
        agg = _weighted_histogram(xs, self.num_opinions, self.weights)[:, self.idx_to_return]
        if self.normalize:
            agg = agg / (xs.shape[1] if self.normalization is None else self.normalization)
        return agg[0] if was_1d else agg

    # -- what the simulation kernels need to compute this CV on the fly -----------------------
    def _kernel_spec(self, num_agents: int):
        """KernelCV when the CV is an integer-weighted histogram (exact in-kernel), else None."""
        from .engine import KernelCV

        idx = self.idx_to_return.astype(np.int32)
        if self.weights is None:
            return KernelCV(idx, float(num_agents) if self.normalize else 1.0)
        w = np.asarray(self.weights, dtype=np.float64)
        if w.shape != (num_agents,) or not np.all(w == np.round(w)) or np.sum(np.abs(w)) >= 2**31:
            return None  # float weights: evaluated from stored states (sponet_opinion_shares)
        return KernelCV(idx, float(self.normalization) if self.normalize else 1.0, agent_weight=w.astype(np.int32))


class DegreeWeightedOpinionShares(OpinionShares):
    """OpinionShares with each agent weighted by its degree (collective_variables.py:127-149)."""

    def __init__(self, num_opinions: int, network, normalize: bool = False, idx_to_return: ArrayLike | None = None):
        weights = np.array([d for _, d in network.degree()])
        super().__init__(num_opinions, normalize, weights, idx_to_return)


class OpinionSharesByDegree:
    """Opinion counts within each degree class, smallest degree first (collective_variables.py:152-220)."""

    def __init__(self, num_opinions: int, network, normalize: bool = False, idx_to_return: ArrayLike | None = None):
        self.degrees_of_nodes = np.array([d for _, d in network.degree()])
        self.degrees = np.sort(np.unique(self.degrees_of_nodes))
        self.num_opinions = num_opinions
        self.normalize = normalize
        self.idx_to_return = _select(idx_to_return, num_opinions)
        self.dimension = len(self.idx_to_return) * self.degrees.shape[0]

    def _kernel_spec(self, num_agents: int):
        """One class per distinct degree; out = counts[class, opinion] (/ class size when normalised)."""
        from .engine import KernelCV

        if self.degrees_of_nodes.shape != (num_agents,):
            return None
        M, C = self.num_opinions, self.degrees.shape[0]
        cls = np.searchsorted(self.degrees, self.degrees_of_nodes).astype(np.int32)
        idx = (np.arange(C)[:, None] * M + self.idx_to_return[None, :]).ravel()
        divisors = None
        if self.normalize:
            sizes = np.bincount(cls, minlength=C).astype(np.float64)
            divisors = np.repeat(sizes, len(self.idx_to_return))
        return KernelCV(idx, 1.0, divisors=divisors, num_classes=C, agent_class=cls)

    def __call__(self, x):
        xs, was_1d = _as_2d(x)
        k = len(self.idx_to_return)
        blocks = []
        for deg in self.degrees:
            mask = (self.degrees_of_nodes == deg).astype(np.float64)
            agg = _weighted_histogram(xs, self.num_opinions, mask)[:, self.idx_to_return]
            if self.normalize:
                agg = agg / np.sum(mask)
            blocks.append(agg)
        if isinstance(blocks[0], np.ndarray):
            cv = np.concatenate(blocks, axis=1)
        else:
            import torch

            cv = torch.cat(blocks, dim=1)
        assert cv.shape[1] == k * len(self.degrees)
        return cv[0] if was_1d else cv


class CompositeCollectiveVariable:
    """Concatenation of several CVs (collective_variables.py:223-256)."""

    def __init__(self, collective_variables: list[CollectiveVariable]):
        self.collective_variables = collective_variables
        self.dimension = sum(cv.dimension for cv in collective_variables)

    def _kernel_spec(self, num_agents: int):
        """Concatenation is in-kernel when every part reads the same (class, weight) table."""
        from .engine import KernelCV

        specs = [cv._kernel_spec(num_agents) if hasattr(cv, "_kernel_spec") else None for cv in self.collective_variables]
        if not specs or any(sp is None or sp.edge for sp in specs):
            return None

        def same(a, b):
            return (a is None and b is None) or (a is not None and b is not None and np.array_equal(a, b))

        first = specs[0]
        for sp in specs[1:]:
            if sp.num_classes != first.num_classes or not same(sp.agent_class, first.agent_class) or not same(
                sp.agent_weight, first.agent_weight
            ):
                return None
        return KernelCV(
            np.concatenate([sp.idx for sp in specs]), 1.0, divisors=np.concatenate([sp.divisor_vector() for sp in specs]),
            num_classes=first.num_classes, agent_class=first.agent_class, agent_weight=first.agent_weight,
        )

    def __call__(self, x):
        xs, was_1d = _as_2d(x)
        parts = [cv(xs) for cv in self.collective_variables]
        if all(isinstance(p, np.ndarray) for p in parts):
            out = np.concatenate(parts, axis=1)
        else:
            import torch

            out = torch.cat([p if hasattr(p, "data_ptr") else torch.as_tensor(p, device=xs.device) for p in parts], dim=1)
        return out[0] if was_1d else out


def _host_states(x):
    xs, was_1d = _as_2d(x)
    if not isinstance(xs, np.ndarray):
        xs = xs.cpu().numpy()
    return np.ascontiguousarray(xs), was_1d


class Interfaces:
    """Number of edges between an agent of opinion 0 and one of opinion 1 (collective_variables.py:259-305).

    Only for two opinions.  ``normalize`` divides by the number of edges.  The count runs on the GPU
    over the CSR of ``network`` (``sponet_interfaces``).
    """

    def __init__(self, network, normalize: bool = False):
        from .engine import GraphCache
        from .utils import calculate_neighbor_list

        self.dimension = 1
        self.normalize = normalize
        self.network = network
        self._neighbor_list = calculate_neighbor_list(network)
        self._graphs = GraphCache()

    def _kernel_spec(self, num_agents: int):
        """In-kernel: the simulation kernels scan this CV's CSR against the on-chip state per output time."""
        from .engine import KernelCV

        if len(self._neighbor_list) != num_agents:
            return None
        div = float(self.network.number_of_edges()) if self.normalize else 1.0
        return KernelCV([], div, kind=KernelCV.INTERFACES, graph=self._graphs.get(self._neighbor_list))

    def __call__(self, x):
        xs, was_1d = _host_states(x)
        if xs.size and np.max(xs) > 1:
            raise ValueError("Interfaces can only be used for 2 opinions.")
        xs = xs.astype(np.uint8, copy=False)
        out = np.empty((xs.shape[0], 1), dtype=np.float64)
        g = self._graphs.get(self._neighbor_list)
        _lib.check(_lib.lib().sponet_interfaces(g.handle, _lib.ptr(xs), xs.shape[0], _lib.ptr(out), None))
        if self.normalize:
            out /= self.network.number_of_edges()
        return out[0] if was_1d else out


class Propensities:
    """Cumulative transition rates (prop_01, prop_10) of a 2-opinion CNVM (collective_variables.py:308-397)."""

    def __init__(self, params, normalize: bool = False):
        from .engine import GraphCache

        self.dimension = 2
        self.params = params
        self.normalize = normalize
        self._graphs = GraphCache()

    def _kernel_spec(self, num_agents: int):
        """In-kernel on a network (the complete-graph form is evaluated from stored states)."""
        from .engine import KernelCV

        p = self.params
        if p.network is None or p.num_agents != num_agents or p.num_opinions != 2:
            return None
        das = np.array([len(nb) for nb in p.network]) ** p.alpha
        r, rt = np.asarray(p.r, dtype=np.float64), np.asarray(p.r_tilde, dtype=np.float64)
        return KernelCV([], float(p.num_agents) if self.normalize else 1.0, kind=KernelCV.PROPENSITIES,
                        graph=self._graphs.get(p.network), degree_alpha=das, rates=(r[0, 1], r[1, 0], rt[0, 1], rt[1, 0]))

    def __call__(self, x):
        xs, was_1d = _host_states(x)
        if xs.size and np.max(xs) > 1:
            raise ValueError("Propensities can only be used for 2 opinions.")
        xs = xs.astype(np.uint8, copy=False)
        p = self.params
        r = np.ascontiguousarray(p.r, dtype=np.float64)
        rt = np.ascontiguousarray(p.r_tilde, dtype=np.float64)
        out = np.empty((xs.shape[0], 2), dtype=np.float64)
        if p.network is None:
            da = float((p.num_agents - 1) ** p.alpha)
            rc = _lib.lib().sponet_propensities(None, p.num_agents, _lib.ptr(xs), xs.shape[0], None, da, _lib.ptr(r),
                                                _lib.ptr(rt), _lib.ptr(out), None)
        else:
            das = np.ascontiguousarray(np.array([len(nb) for nb in p.network]) ** p.alpha, dtype=np.float64)
            g = self._graphs.get(p.network)
            rc = _lib.lib().sponet_propensities(g.handle, p.num_agents, _lib.ptr(xs), xs.shape[0], _lib.ptr(das), 0.0,
                                                _lib.ptr(r), _lib.ptr(rt), _lib.ptr(out), None)
        _lib.check(rc)
        if self.normalize:
            out /= p.num_agents
        return out[0] if was_1d else out
