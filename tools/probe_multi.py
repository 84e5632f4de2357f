"""Development aid: device-timed throughput of the CNVM kernel for a list of (n, runs, t_max, T, warps_per_chain)
configurations in one process (one graph build per n).  SPONET_B200_LIB selects the library build.
    python tools/probe_multi.py "100000:10000:10:101:0" "100000:1184:2:11:2" ...
"""
import os
import sys
import time

import numpy as np
import torch

from sponet_b200 import CNTM, CNVM, engine, workloads

torch.cuda.init()
models = {}
print("lib:", os.environ.get("SPONET_B200_LIB", "default"))
for spec_s in sys.argv[1:]:
    parts = spec_s.split(":")
    kind = "h"
    if not parts[0].isdigit():  # optional workload prefix: h (headline ER, M=3) | sis (WS ring, M=2)
        kind, parts = parts[0], parts[1:]
    glob = len(parts) > 5 and parts[5] == "g"
    n, runs, t_max, T, wpc = parts[:5]
    n, runs, T, wpc, t_max = int(n), int(runs), int(T), int(wpc), float(t_max)
    if (kind, n) not in models:
        if kind == "cntm":
            params, x0, rate = workloads.ba_cntm(n)
            model = CNTM(params)
            models[(kind, n)] = (params, x0, rate, model, (model._struct(), None), model._graph())
        else:
            params, x0, rate = workloads.headline_cnvm(n) if kind == "h" else workloads.sis_ws(n)
            model = CNVM(params)
            models[(kind, n)] = (params, x0, rate, model, model._struct(), model._graph())
    params, x0, rate, model, (struct, keep), graph = models[(kind, n)]
    M = 2 if kind == "cntm" else params.num_opinions
    ek = "cntm" if kind == "cntm" else "cnvm"
    kw = {} if kind == "cntm" else dict(warps_per_chain=wpc)
    t_eval = np.linspace(0, t_max, T)
    d_x0 = torch.as_tensor(x0[None, :], device="cuda")
    spec = (np.arange(M, dtype=np.int32), 1.0)
    tw = time.time()
    while time.time() - tw < 1.0:
        engine.run_native(ek, graph, n, struct, d_x0, t_eval, 1, runs, 0, runs, cv_spec=spec, want_cv=True, force_global_state=glob, **kw)
        torch.cuda.synchronize()
    best = None
    for rep in range(3):
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        engine.run_native(ek, graph, n, struct, d_x0, t_eval, 1234 + rep, runs, 0, runs, cv_spec=spec, want_cv=True,
                          counters=counters, force_global_state=glob, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c = counters.cpu().numpy()
        r = (c[0] / ms * 1e3, ms, c[3] / max(c[2], 1))
        best = r if best is None or r[0] > best[0] else best
    print(f"{kind}{" global" if glob else ""} n={n} runs={runs} t_max={t_max} T={T} wpc={wpc}: {best[0]:.4g} ev/s  ({best[1]:.1f} ms)  per-chain {best[0]/runs:.4g}  extra rounds/step {best[2]:.4f}", flush=True)
