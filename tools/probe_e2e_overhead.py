"""Fixed host-side cost of a sample_many_runs call on H (t_max -> 0): python tools/probe_e2e_overhead.py"""
import cProfile
import pstats
import time

import numpy as np

from sponet_b200 import OpinionShares, sample_many_runs, workloads

params, x0, _ = workloads.headline_cnvm(100_000)
cv = OpinionShares(3)
for t_max in (0.01, 0.01, 1.0):
    t0 = time.perf_counter()
    sample_many_runs(params, x0, t_max, 101, 10_000, collective_variable=cv, rng=np.random.default_rng(1))
    print(f"t_max={t_max}: {1e3 * (time.perf_counter() - t0):.1f} ms")
pr = cProfile.Profile()
pr.enable()
sample_many_runs(params, x0, 0.01, 101, 10_000, collective_variable=cv, rng=np.random.default_rng(1))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
