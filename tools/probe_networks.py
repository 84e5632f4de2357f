"""Per-run networks (a NetworkGenerator in the parameters) at the H shape: device-generated batches against the host pipeline.
    python tools/probe_networks.py [R] [t_max]"""
import sys
import time

import numpy as np
import torch

from sponet_b200 import CNVMParameters, OpinionShares, engine, sample_many_runs
from sponet_b200 import multiprocessing as mp
from sponet_b200.network_generator import ErdosRenyiGenerator

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
t_max = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
N, M = 100_000, 3
p = 10 / (N - 1)
kw = dict(num_opinions=M, r=[[0, 1, 2], [1, 0, 1], [2, 0, 0]], r_tilde=[[0, .2, .1], [0, 0, .1], [.1, .2, 0]])
x0 = np.random.default_rng(0).integers(0, M, N)
cv = OpinionShares(M)
events = R * 2.6e5 * t_max

for n_graphs in (148, 1184, 2368):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = engine.GraphBatch(N, p, 5, 1, n_graphs, 0, n_graphs, force_no_isolates=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = b.info()
    print(f"generate {n_graphs} graphs G({N}, {p:.2e}): {dt * 1e3:.1f} ms = {dt / n_graphs * 1e6:.1f} us per graph, nnz {info['nnz']}, degrees {info['min_degree']}..{info['max_degree']}")
    b.close()

# the simulation kernel alone on a resident batch (CUDA events)
from sponet_b200 import CNVM

model = CNVM(CNVMParameters(num_agents=N, **kw))
struct, keep = model._struct()
spec = (np.arange(M, dtype=np.int32), 1.0)
d_x0 = torch.as_tensor(x0[None, :].astype(np.uint8), device="cuda")
for n_graphs in (1184, min(R, 2368)):
    b = engine.GraphBatch(N, p, 5, 1, n_graphs, 0, n_graphs, force_no_isolates=True)
    t_eval = np.linspace(0, t_max, 11)
    best = 0.0
    for rep in range(4):
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        engine.run_native("cnvm", b, N, struct, d_x0, t_eval, 77 + rep, n_graphs, 0, n_graphs, cv_spec=spec, want_cv=True, counters=counters)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, counters[0].item() / e0.elapsed_time(e1) * 1e3)
    print(f"kernel on a batch of {n_graphs} graphs, t_max={t_max}: {best:.4g} events/s")
    b.close()


def params():
    return CNVMParameters(network_generator=ErdosRenyiGenerator(N, p, rng=np.random.default_rng(1), force_no_isolates=True), **kw)


for rep in range(3):
    t0 = time.perf_counter()
    _, c = sample_many_runs(params(), x0, t_max, 11, R, collective_variable=cv, rng=np.random.default_rng(2))
    dt = time.perf_counter() - t0
    print(f"device batches: R={R} t_max={t_max}: {dt:.3f} s, {events / dt:.3e} events/s  (mean final shares {c[:, -1].mean(0) / N})")

mp._FORCE_HOST_NETWORKS = True
Rh = 8
t0 = time.perf_counter()
_, c = sample_many_runs(params(), x0, t_max, 11, Rh, collective_variable=cv, rng=np.random.default_rng(2))
dt = time.perf_counter() - t0
print(f"host pipeline (networkx per run): R={Rh}: {dt:.3f} s = {dt / Rh:.3f} s per run, {Rh * 2.6e5 * t_max / dt:.3e} events/s")
