"""CNVM front-end (reference: sponet/cnvm/model.py:12-181).

``CNVM(params).simulate(t_max, x_init, t_eval, rng)`` keeps the reference signature and return
shapes; the event loop itself runs on the GPU behind the C ABI:

* ``mode="native"`` (default): Philox-driven warp-per-chain kernel (``sponet_cnvm_run``); ``rng``
  supplies the 64-bit key, so equal seeds give equal results.  Returned times are the ideal grid
  ``t_eval`` (native mode snapshots the state AT ``t_eval[k]``).
* ``mode="replay"``: the reference loop re-executed draw for draw on ``rng``'s raw uint64 stream
  (``sponet_cnvm_replay``): states and event times are bit-identical to the reference, and ``rng`` is
  left advanced exactly as the reference would leave it.
* ``t_eval=None`` (store every change) and ``num_opinions > 256`` always take the replay kernels.
"""
from __future__ import annotations

import numpy as np
from numpy.random import Generator, default_rng
from numpy.typing import ArrayLike, NDArray

from .. import engine
from ..utils import argmatch, t_eval_to_ndarray
from .parameters import CNVMParameters


class CNVM:
    def __init__(self, params: CNVMParameters):
        self.params = params
        self._graphs = engine.GraphCache()
        self._calculate_degree_alpha()

    # reference: sponet/cnvm/model.py:30-42
    def _calculate_degree_alpha(self) -> None:
        p = self.params
        if p.network is not None:
            degrees = np.array([len(nbrs) for nbrs in p.network])
            if np.min(degrees) < 1:
                raise ValueError("Isolated vertices in the network are not allowed.")
            self.degree_alpha = degrees ** (1 - p.alpha)
        else:
            self.degree_alpha = np.ones(p.num_agents) * (p.num_agents - 1) ** (1 - p.alpha)

    def update_network(self) -> None:
        """Resample the network from the NetworkGenerator in params (reference :44-50)."""
        self.params.update_network_by_generator()
        self._graphs.clear()
        self._calculate_degree_alpha()

    def update_rates(self, r: ArrayLike | None = None, r_tilde: ArrayLike | None = None) -> None:
        self.params.change_rates(r, r_tilde)

    # -- kernel-facing pieces ---------------------------------------------------------------
    @property
    def is_complete(self) -> bool:
        return self.params.network is None

    def _graph(self):
        return None if self.is_complete else self._graphs.get(self.params.network)

    def _struct(self):
        return engine.cnvm_struct(self.params, self.degree_alpha)

    def event_rates(self) -> tuple[float, float]:
        """(next_event_rate, noise_probability) as the reference loops compute them."""
        return engine.cnvm_event_rates(self.params, self.degree_alpha, self.is_complete)

    def _draws_per_event(self) -> float:
        # network loop: branch, {agent, opinion, accept | alias, slot, accept}, ziggurat (~1.02)
        # complete loop: agent, branch, {opinion, accept | neighbour (rejection), accept}, ziggurat
        return 5.05 if not self.is_complete else 5.05 + 2.0 / max(self.params.num_agents, 2)

    # -- public -------------------------------------------------------------------------------
    def simulate(
        self,
        t_max: float,
        x_init: ArrayLike | None = None,
        t_eval: ArrayLike | None = None,
        rng: Generator = default_rng(),
        mode: str | None = None,
    ) -> tuple[NDArray, NDArray]:
        """Simulate from t=0 to ``t_max``; returns ``(t_traj, x_traj)`` with ``x_traj.shape == (len(t_traj), N)``."""
        p = self.params
        if p.network_generator is not None:
            self.update_network()
        dtype = np.min_scalar_type(p.num_opinions - 1)
        if x_init is None:
            x = rng.choice(np.arange(p.num_opinions, dtype=dtype), size=p.num_agents)
        else:
            xi = np.asarray(x_init)
            if xi.size and (xi.min() < 0 or xi.max() >= p.num_opinions):
                raise ValueError("x_init must hold opinions in [0, num_opinions)")
            x = np.array(xi, dtype=dtype)
        if x.shape != (p.num_agents,):
            raise ValueError(f"x_init must have shape ({p.num_agents},)")
        if t_eval is not None:
            t_eval = t_eval_to_ndarray(t_eval, t_max)
        mode = engine.resolve_mode(mode)
        struct, keep = self._struct()
        rate, _ = self.event_rates()

        if t_eval is None:
            t_traj, x_traj = engine.run_replay_all(
                "cnvm", self._graph(), p.num_agents, struct, x, t_max, rng, t_max / rate, self._draws_per_event()
            )
        elif mode == "replay" or p.num_opinions > 256:
            budget = engine.stream_budget((t_eval[-1] - t_eval[0]) / rate + 1, self._draws_per_event())
            out_t, out_x, _ = engine.run_replay(
                "cnvm", self._graph(), p.num_agents, struct, x[None, :], 1, t_eval, [rng], [(0, 1, 0, 1)], budget
            )
            t_traj, x_traj = out_t[0, 0], out_x[0, 0]
        else:
            out_x, _ = engine.run_native(
                "cnvm", self._graph(), p.num_agents, struct, np.ascontiguousarray(x[None, :]), t_eval,
                engine.draw_seed(rng), 1, 0, 1, want_x=True,
            )
            t_traj, x_traj = t_eval.copy(), out_x[0, 0].astype(dtype, copy=False)
        del keep

        if t_eval is not None and t_traj.shape[0] != t_eval.shape[0]:  # reference :113-118
            t_ind = argmatch(t_eval, t_traj)
            t_traj, x_traj = t_eval, x_traj[t_ind]
        return t_traj, x_traj
