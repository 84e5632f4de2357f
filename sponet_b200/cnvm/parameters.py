"""CNVM parameter container (reference: sponet/cnvm/parameters.py).

Same constructor, attributes, precedence rules and ValueError cases as the reference class; the
arithmetic that feeds the kernels (``convert_rate_to_cnvm``) is restated here in numpy float64.

Rate model: agent i with opinion m switches to n != m at rate
``r[m, n] * d(i, n) / d(i)**alpha + r_tilde[m, n]``, equivalently
``r_imit * d(i, n) / d(i)**alpha * prob_imit[m, n] + r_noise / M * prob_noise[m, n]``.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np
from numpy.typing import ArrayLike, NDArray

from ..utils import calculate_neighbor_list


def _is_graph(obj) -> bool:
    return hasattr(obj, "number_of_nodes") and hasattr(obj, "neighbors")


class CNVMParameters:
    """Parameters of the continuous-time noisy voter model.

    Exactly one source of the network is used, in this order of precedence (reference
    sponet/cnvm/parameters.py:150-166): ``network_generator`` > ``network`` > ``num_agents``
    (complete graph).  ``network`` is an ``nx.Graph`` or an adjacency list.

    Rates come in "style 1" (``r``, ``r_tilde``: matrices or scalars) or "style 2"
    (``r_imit``, ``r_noise``, ``prob_imit``, ``prob_noise``).
    """

    def __init__(
        self,
        num_opinions: int | None = None,
        num_agents: int | None = None,
        network=None,
        network_generator=None,
        alpha: float = 1.0,
        r: ArrayLike | None = None,
        r_tilde: ArrayLike | None = None,
        r_imit: float | None = None,
        r_noise: float | None = None,
        prob_imit: ArrayLike = 1,
        prob_noise: ArrayLike = 1,
    ):
        self.alpha = alpha
        self.num_agents, self.network, self.network_generator = _resolve_network(
            num_agents, network, network_generator
        )
        self._set_rates(_resolve_rates(num_opinions, r, r_tilde, r_imit, r_noise, prob_imit, prob_noise))

    def _set_rates(self, resolved) -> None:
        (
            self.num_opinions,
            self.r,
            self.r_tilde,
            self.r_imit,
            self.r_noise,
            self.prob_imit,
            self.prob_noise,
        ) = resolved

    # -- network -----------------------------------------------------------------------------
    def get_network(self):
        """The network as an ``nx.Graph`` (generated / complete graph if none is stored)."""
        import networkx as nx

        if self.network is not None:
            return nx.from_dict_of_lists(dict(enumerate(self.network)))
        if self.network_generator is not None:
            return self.network_generator()
        return nx.complete_graph(self.num_agents)

    def set_network(self, network) -> None:
        if self.network_generator is not None:
            raise ValueError("Cannot set network when there is a NetworkGenerator.")
        self.num_agents, self.network, _ = _resolve_network(self.num_agents, network, None)

    def update_network_by_generator(self) -> None:
        if self.network_generator is None:
            raise ValueError("No NetworkGenerator was given.")
        self.network = calculate_neighbor_list(self.network_generator())

    # -- rates -------------------------------------------------------------------------------
    def change_rates(self, r: ArrayLike | None = None, r_tilde: ArrayLike | None = None) -> None:
        """Replace ``r`` and/or ``r_tilde``; the omitted one keeps its value."""
        r = self.r if r is None else r
        r_tilde = self.r_tilde if r_tilde is None else r_tilde
        self._set_rates(_resolve_rates(self.num_opinions, r, r_tilde, None, None, 1, 1))


def _resolve_network(num_agents, network, network_generator):
    if network_generator is not None:
        return network_generator.num_agents, None, network_generator
    if network is not None:
        if _is_graph(network):
            return network.number_of_nodes(), calculate_neighbor_list(network), None
        if isinstance(network, list) and len(network) and isinstance(network[0], np.ndarray):
            adjacency = network
        else:
            adjacency = [np.array(nbrs, dtype=np.int32) for nbrs in network]
        return len(adjacency), adjacency, None
    if num_agents is not None:
        return num_agents, None, None
    raise ValueError("Either a network or a NetworkGenerator or num_agents has to be specified.")


def _resolve_rates(num_opinions, r, r_tilde, r_imit, r_noise, prob_imit, prob_noise):
    if r is not None and r_tilde is not None:  # style 1
        r, r_tilde, num_opinions = _as_rate_matrices(num_opinions, r, r_tilde)
        if np.min(r) < 0 or np.min(r_tilde) < 0:
            raise ValueError("Rates have to be non-negative.")
        return (num_opinions, r, r_tilde, *convert_rate_to_cnvm(r, r_tilde))
    if r_imit is not None and r_noise is not None:  # style 2
        prob_imit, prob_noise, num_opinions = _as_rate_matrices(num_opinions, prob_imit, prob_noise)
        if r_imit < 0 or r_noise < 0:
            raise ValueError("Rates have to be non-negative.")
        for p in (prob_imit, prob_noise):
            if np.min(p) < 0 or np.max(p) > 1:
                raise ValueError("Probabilities have to be between 0 and 1.")
        return (
            num_opinions,
            r_imit * prob_imit,
            r_noise * prob_noise / num_opinions,
            r_imit,
            r_noise,
            prob_imit,
            prob_noise,
        )
    raise ValueError("Rate parameters have to be provided.")


def _as_rate_matrices(num_opinions, a, b):
    """Scalars become constant off-diagonal matrices; matrices get a zero diagonal; shapes are checked."""
    scalar_a, scalar_b = isinstance(a, (int, float)), isinstance(b, (int, float))
    if scalar_a and scalar_b:
        if num_opinions is None:
            raise ValueError("`num_opinions` has to be provided.")
        ones = np.ones((num_opinions, num_opinions))
        a, b = a * ones, b * ones
        np.fill_diagonal(a, 0)
        np.fill_diagonal(b, 0)
        return a, b, num_opinions
    if scalar_a:
        b = np.array(b, ndmin=2)
        a = a * np.ones_like(b)
    elif scalar_b:
        a = np.array(a, ndmin=2)
        b = b * np.ones_like(a)
    else:
        a, b = np.array(a, ndmin=2), np.array(b, ndmin=2)
    np.fill_diagonal(a, 0)
    np.fill_diagonal(b, 0)
    for mat in (a, b):
        if mat.ndim != 2 or mat.shape[0] != mat.shape[1]:
            raise ValueError("Expected square matrix.")
    if a.shape[0] != b.shape[0]:
        raise ValueError("Shapes do not match.")
    if num_opinions is not None and (a.shape[0] != num_opinions or b.shape[0] != num_opinions):
        raise ValueError("Shape of matrix does not match `num_opinions`")
    return a, b, int(a.shape[0])


def convert_rate_to_cnvm(r: ArrayLike, r_tilde: ArrayLike) -> tuple[float, float, NDArray, NDArray]:
    """``(r, r_tilde) -> (r_imit, r_noise, prob_imit, prob_noise)`` (sponet/cnvm/parameters.py:293-338).

    ``r_imit = max offdiag r``, ``prob_imit = r / r_imit``; ``r_noise = M * max offdiag r_tilde``,
    ``prob_noise = r_tilde * M / r_noise``; all-zero rates give zero probability matrices.
    """
    r, r_tilde = np.array(r), np.array(r_tilde)
    M = r.shape[0]

    def split(mat, scale):
        off = np.copy(mat)
        np.fill_diagonal(off, 0)
        rate = np.max(off) * scale
        prob = off * scale / rate if rate > 0 else np.zeros((M, M))
        return rate, prob

    r_imit, prob_imit = split(r, 1)
    r_noise, prob_noise = split(r_tilde, M)
    return r_imit, r_noise, prob_imit, prob_noise


def convert_rate_from_cnvm(params: CNVMParameters) -> tuple[NDArray, NDArray]:
    """Inverse of :func:`convert_rate_to_cnvm` (sponet/cnvm/parameters.py:341-361)."""
    return params.r_imit * params.prob_imit, params.r_noise * params.prob_noise / params.num_opinions


def save_params_as_txt_file(filename: str, params: CNVMParameters) -> None:
    """Human-readable dump (sponet/cnvm/parameters.py:262-290)."""
    path = filename if filename.endswith(".txt") else filename + ".txt"
    if params.network_generator is not None:
        net = f"network_generator = {params.network_generator}"
    elif params.network is not None:
        net = f"network = graph with {len(params.network)} nodes"
    else:
        net = "network = fully connected"
    with open(path, "w") as fh:
        fh.write(
            f"num_agents = {params.num_agents}\nnum_opinions = {params.num_opinions}\n"
            f"alpha = {params.alpha}\nr_imit = {params.r_imit}\nr_noise = {params.r_noise}\n\n"
            f"prob_imit =\n {params.prob_imit}\n\nprob_noise =\n {params.prob_noise}\n\n{net}\n"
        )
