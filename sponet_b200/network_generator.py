"""Network generators for ``params.network_generator`` (a new network for every run).

Mirrors the call surface of the reference's generators that SURVEY.md 8(f) names (sponet/network_generator.py:26-183:
Erdos-Renyi, random regular, Barabasi-Albert, Watts-Strogatz): constructor arguments, ``num_agents``, ``__call__() ->
nx.Graph`` on the host through networkx, ``__repr__`` and ``abrv()``.  Any other object with that protocol (the
reference's own generators included) works as well.

What is new: a generator may also offer ``device_batch(...)``, which makes ALL the graphs of an ensemble call on the
GPU (``engine.GraphBatch``), one per realization, so that ``sample_many_runs`` runs the whole ensemble in one launch
instead of one host-generated graph and one launch per realization.  All four generators
of this module implement it (``csrc/graphgen.cu``).  The device graphs follow the same random-graph process as
networkx's, but they are not the same graphs: they are keyed by (seed, realization index), the seed being drawn from the
generator's ``rng``.
"""
from __future__ import annotations

import time
from typing import Protocol

import numpy as np
from numpy.random import Generator, default_rng


class NetworkGenerator(Protocol):
    num_agents: int

    def __call__(self): ...

    def __repr__(self) -> str: ...

    def abrv(self) -> str: ...


def _nx():
    import networkx as nx

    return nx


def _connect_isolates(network, rng: Generator) -> None:
    """One edge to a uniformly drawn other vertex for every vertex that is (still) isolated, in vertex order
    (reference :446-454)."""
    n = network.number_of_nodes()
    for v in list(network.nodes):
        if network.degree(v) > 0:
            continue
        w = v
        while w == v:
            w = int(rng.integers(0, n))
        network.add_edge(v, w)


class ErdosRenyiGenerator:
    """Binomial random graphs G(num_agents, p) without isolated vertices (reference :26-85).

    A sample with an isolated vertex is drawn again until ``max_sample_time`` seconds have passed (``RuntimeError``),
    or, with ``force_no_isolates``, repaired by one random edge per isolated vertex."""

    # device batches give up after this many rejected samples per graph (the host path has its time limit)
    max_device_attempts = 64

    def __init__(self, num_agents: int, p: float, max_sample_time: float = 10, rng: Generator = default_rng(),
                 force_no_isolates: bool = False):
        self.num_agents = num_agents
        self.p = p
        self.max_sample_time = max_sample_time
        self.rng = rng
        self.force_no_isolates = force_no_isolates

    def __call__(self):
        nx = _nx()
        sampler = nx.erdos_renyi_graph if self.p > 0.2 else nx.fast_gnp_random_graph
        deadline = time.time() + self.max_sample_time
        while True:
            network = sampler(self.num_agents, self.p, seed=self.rng)
            if nx.number_of_isolates(network) == 0:
                return network
            if self.force_no_isolates:
                _connect_isolates(network, self.rng)
                return network
            if time.time() > deadline:
                raise RuntimeError("Timeout. Could not generate a graph without isolated vertices.")

    def draw_batch_seed(self) -> int:
        """Advances ``rng`` once per ensemble call; every graph of the call derives from this seed and its index."""
        return int(self.rng.integers(0, 2**63, dtype=np.int64))

    def expected_mean_degree(self) -> float:
        return self.p * (self.num_agents - 1)

    def device_batch(self, seed: int, num_states: int, num_runs_total: int, run_begin: int, run_end: int,
                     padded_rows: bool = False, records: bool = True):
        """The graphs of realizations [run_begin, run_end) of every initial state, generated on the current GPU."""
        from . import engine

        return engine.GraphBatch(self.num_agents, self.p, seed, num_states, num_runs_total, run_begin, run_end,
                                 force_no_isolates=self.force_no_isolates, max_attempts=self.max_device_attempts,
                                 padded_rows=padded_rows, records=records)

    def __repr__(self) -> str:
        return f"Erdos-Renyi random graph with p={self.p} on {self.num_agents} nodes"

    def abrv(self) -> str:
        return f"ER_p{int(self.p * 100)}_N{self.num_agents}"


class RandomRegularGenerator:
    """Uniform random d-regular graphs (reference :87-115)."""

    def __init__(self, num_agents: int, d: int, rng: Generator = default_rng()):
        self.num_agents = num_agents
        self.d = d
        self.rng = rng
        self.build_bytes_per_vertex = 8.0 * d + 4.0  # stubs and adjacency rows while a device batch is built

    def __call__(self):
        return _nx().random_regular_graph(self.d, self.num_agents, seed=self.rng)

    def draw_batch_seed(self) -> int:
        return int(self.rng.integers(0, 2**63, dtype=np.int64))

    def expected_mean_degree(self) -> float:
        return float(self.d)

    def device_batch(self, seed: int, num_states: int, num_runs_total: int, run_begin: int, run_end: int,
                     padded_rows: bool = False, records: bool = True):
        """networkx's pairing process with restarts, run on the current GPU (one warp per graph) for the realizations
        [run_begin, run_end) of every initial state."""
        from . import engine

        return engine.GraphBatch.random_regular(self.num_agents, self.d, seed, num_states, num_runs_total, run_begin, run_end,
                                                padded_rows=padded_rows, records=records)

    def __repr__(self) -> str:
        return f"Uniform random regular graph with d={self.d} on {self.num_agents} nodes"

    def abrv(self) -> str:
        return f"regular_d{self.d}_N{self.num_agents}"


class BarabasiAlbertGenerator:
    """Preferential attachment with m edges per new vertex (reference :118-144)."""

    def __init__(self, num_agents: int, m: int, rng: Generator = default_rng()):
        self.num_agents = num_agents
        self.m = m
        self.rng = rng
        self.build_bytes_per_vertex = 4.0 * m  # the targets of every vertex while a device batch is built

    def __call__(self):
        return _nx().barabasi_albert_graph(self.num_agents, self.m, seed=self.rng)

    def draw_batch_seed(self) -> int:
        return int(self.rng.integers(0, 2**63, dtype=np.int64))

    def expected_mean_degree(self) -> float:
        return 2.0 * self.m

    def device_batch(self, seed: int, num_states: int, num_runs_total: int, run_begin: int, run_end: int,
                     padded_rows: bool = False, records: bool = True):
        """The same process as networkx's (star on m + 1 vertices, m distinct degree-proportional targets per new vertex),
        run on the current GPU for the realizations [run_begin, run_end) of every initial state."""
        from . import engine

        return engine.GraphBatch.barabasi_albert(self.num_agents, self.m, seed, num_states, num_runs_total, run_begin, run_end,
                                                 padded_rows=padded_rows, records=records)

    def __repr__(self) -> str:
        return f"Barabasi-Albert random graph on {self.num_agents} nodes"

    def abrv(self) -> str:
        return f"barabasi_m{self.m}_N{self.num_agents}"


class WattsStrogatzGenerator:
    """Connected small-world graphs: ring of ``num_neighbors`` nearest neighbours, rewired with probability p
    (reference :147-183)."""

    def __init__(self, num_agents: int, num_neighbors: int, p: float, rng: Generator = default_rng()):
        self.num_agents = num_agents
        self.num_neighbors = num_neighbors
        self.p = p
        self.rng = rng
        self.build_bytes_per_vertex = 4.0 * (num_neighbors + 32) + 8.0  # adjacency rows under construction, queue, marks

    def __call__(self):
        return _nx().connected_watts_strogatz_graph(self.num_agents, self.num_neighbors, self.p, seed=self.rng)

    def draw_batch_seed(self) -> int:
        return int(self.rng.integers(0, 2**63, dtype=np.int64))

    def expected_mean_degree(self) -> float:
        return 2.0 * (self.num_neighbors // 2)

    def device_batch(self, seed: int, num_states: int, num_runs_total: int, run_begin: int, run_end: int,
                     padded_rows: bool = False, records: bool = True):
        """networkx's process (ring lattice, sequential rewiring against the current edges, drawn again until connected) run
        on the current GPU, one warp per graph, for the realizations [run_begin, run_end) of every initial state."""
        from . import engine

        return engine.GraphBatch.watts_strogatz(self.num_agents, self.num_neighbors, self.p, seed, num_states, num_runs_total,
                                                run_begin, run_end, padded_rows=padded_rows, records=records)

    def __repr__(self) -> str:
        return f"Watts-Strogatz random graph on {self.num_agents} nodes"

    def abrv(self) -> str:
        return f"watts_k{self.num_neighbors}_p{int(self.p * 100)}_N{self.num_agents}"
