"""CNTM front-end (reference: sponet/cntm/model.py:17-108).

Same ``simulate`` contract as :class:`sponet_b200.cnvm.CNVM`; the loops are ``sponet_cntm_run`` (native)
and ``sponet_cntm_replay`` / ``sponet_cntm_replay_all`` (bit-exact replay).
"""
from __future__ import annotations

import numpy as np
from numpy.random import Generator, default_rng
from numpy.typing import ArrayLike, NDArray

from .. import engine
from ..utils import argmatch, calculate_neighbor_list, t_eval_to_ndarray
from .parameters import CNTMParameters


class CNTM:
    def __init__(self, params: CNTMParameters):
        self.params = params
        self.neighbor_list = calculate_neighbor_list(params.network)
        # reference :31-34
        self.noise_prob = params.r_tilde / (params.r + params.r_tilde)
        self.next_event_rate = 1 / (params.num_agents * (params.r + params.r_tilde))
        self._graphs = engine.GraphCache()

    def _graph(self):
        return self._graphs.get(self.neighbor_list)

    def _struct(self):
        return engine.cntm_struct(self.next_event_rate, self.noise_prob, self.params)

    def _draws_per_event(self) -> float:
        # agent, branch, [random bit if noise], ziggurat (~1.02)
        return 3.05 + self.noise_prob

    def simulate(
        self,
        t_max: float,
        x_init: ArrayLike | None = None,
        t_eval: ArrayLike | None = None,
        rng: Generator = default_rng(),
        mode: str | None = None,
    ) -> tuple[NDArray, NDArray]:
        p = self.params
        dtype = np.min_scalar_type(1)
        if x_init is None:
            x = rng.choice(np.arange(p.num_opinions, dtype=dtype), size=p.num_agents)
        else:
            xi = np.asarray(x_init)
            if xi.size and (xi.min() < 0 or xi.max() >= 2):
                raise ValueError("x_init must hold opinions in [0, num_opinions)")
            x = np.array(xi, dtype=dtype)
        if x.shape != (p.num_agents,):
            raise ValueError(f"x_init must have shape ({p.num_agents},)")
        if t_eval is not None:
            t_eval = t_eval_to_ndarray(t_eval, t_max)
        mode = engine.resolve_mode(mode)
        struct = self._struct()

        if t_eval is None:
            t_traj, x_traj = engine.run_replay_all(
                "cntm", self._graph(), p.num_agents, struct, x, t_max, rng, t_max / self.next_event_rate,
                self._draws_per_event(),
            )
        elif mode == "replay":
            budget = engine.stream_budget((t_eval[-1] - t_eval[0]) / self.next_event_rate + 1, self._draws_per_event())
            out_t, out_x, _ = engine.run_replay(
                "cntm", self._graph(), p.num_agents, struct, x[None, :], 1, t_eval, [rng], [(0, 1, 0, 1)], budget
            )
            t_traj, x_traj = out_t[0, 0], out_x[0, 0]
        else:
            out_x, _ = engine.run_native(
                "cntm", self._graph(), p.num_agents, struct, np.ascontiguousarray(x[None, :]), t_eval,
                engine.draw_seed(rng), 1, 0, 1, want_x=True,
            )
            t_traj, x_traj = t_eval.copy(), out_x[0, 0]

        if t_eval is not None and t_traj.shape[0] != t_eval.shape[0]:
            t_ind = argmatch(t_eval, t_traj)
            t_traj, x_traj = t_eval, x_traj[t_ind]
        return t_traj, x_traj
