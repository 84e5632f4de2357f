"""The CNVM kernel on a resident batch of per-run graphs (ncu target): python tools/probe_batch_kernel.py [graphs] [t_max] [records]"""
import sys

import numpy as np
import torch

from sponet_b200 import CNVM, CNVMParameters, engine

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
t_max = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
records = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
wpc = int(sys.argv[4]) if len(sys.argv) > 4 else 0
import os
N, M = int(os.environ.get("PROBE_N", "100000")), 3
kw = dict(num_opinions=M, r=[[0, 1, 2], [1, 0, 1], [2, 0, 0]], r_tilde=[[0, .2, .1], [0, 0, .1], [.1, .2, 0]])
model = CNVM(CNVMParameters(num_agents=N, **kw))
struct, keep = model._struct()
spec = (np.arange(M, dtype=np.int32), 1.0)
x0 = np.random.default_rng(0).integers(0, M, N).astype(np.uint8)
d_x0 = torch.as_tensor(x0[None, :], device="cuda")
b = engine.GraphBatch(N, 10 / (N - 1), 5, 1, G, 0, G, force_no_isolates=True, records=records)
t_eval = np.linspace(0, t_max, 11)
for rep in range(4):
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    engine.run_native("cnvm", b, N, struct, d_x0, t_eval, 77 + rep, G, 0, G, cv_spec=spec, want_cv=True, counters=counters, warps_per_chain=wpc)
    e1.record()
    torch.cuda.synchronize()
    print(f"N={N} batch of {G} graphs records={records} wpc={wpc} t_max={t_max} extra rounds/step {counters[3].item() / max(counters[2].item(), 1):.3f}: {counters[0].item() / e0.elapsed_time(e1) * 1e3:.4g} events/s ({e0.elapsed_time(e1):.1f} ms)")
