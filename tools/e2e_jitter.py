"""Looks for sporadic slow calls of the public API on workload H (development aid): 30 calls with the Python
garbage collector as it is, 30 after gc.freeze()."""
import gc
import time

import numpy as np
import torch

from sponet_b200 import OpinionShares, sample_many_runs, workloads

params, x0, rate = workloads.headline_cnvm(100_000)
cv = OpinionShares(3)
rng = np.random.default_rng(7)
for _ in range(3):
    sample_many_runs(params, x0, 10.0, 101, 10_000, collective_variable=cv, rng=rng)


def run(label, n=30):
    ts = []
    for k in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sample_many_runs(params, x0, 10.0, 101, 10_000, collective_variable=cv, rng=rng)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = np.array(ts)
    print(f"{label}: median {np.median(ts):.1f} ms, max {ts.max():.1f} ms, calls > 1.1 x median: {(ts > 1.1 * np.median(ts)).sum()}  gc counts {gc.get_count()}", flush=True)
    print("   ", np.round(ts, 0).astype(int).tolist(), flush=True)


run("gc as is")
gc.collect()
gc.freeze()
run("after gc.freeze()")
