"""Host-side overhead of the public API on workload H: per-call wall time of sample_many_runs against the
device-timed kernel (development aid)."""
import time

import numpy as np
import torch

from sponet_b200 import OpinionShares, sample_many_runs, workloads

params, x0, rate = workloads.headline_cnvm(100_000)
cv = OpinionShares(3)
rng = np.random.default_rng(7)
sample_many_runs(params, x0, 10.0, 101, 10_000, collective_variable=cv, rng=rng)
for k in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, x = sample_many_runs(params, x0, 10.0, 101, 10_000, collective_variable=cv, rng=rng)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"call {k}: {dt * 1e3:.1f} ms  -> {10_000 * 10.0 * rate / dt:.4g} events/s", flush=True)

# where the time goes: the ctypes call itself vs everything around it
from sponet_b200 import engine  # noqa: E402

_orig = engine.run_native
spent = []


def timed(*a, **k):
    t0 = time.perf_counter()
    out = _orig(*a, **k)
    spent.append(time.perf_counter() - t0)
    return out


engine.run_native = timed
import sponet_b200.multiprocessing as smp  # noqa: E402

for k in range(6):
    spent.clear()
    t0 = time.perf_counter()
    _, x = sample_many_runs(params, x0, 10.0, 101, 10_000, collective_variable=cv, rng=rng)
    dt = time.perf_counter() - t0
    print(f"call {k}: total {dt * 1e3:.1f} ms, run_native {sum(spent) * 1e3:.1f} ms", flush=True)
