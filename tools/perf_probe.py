"""Quick device-timed throughput probe of the native kernels (development aid, not the bench)."""
import argparse
import os
import time

import numpy as np
import torch

from sponet_b200 import CNVM, CNTM, engine, workloads

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="cnvm")
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--runs", type=int, default=10_000)
ap.add_argument("--t-max", type=float, default=1.0)
ap.add_argument("--T", type=int, default=11)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--global-state", action="store_true")
ap.add_argument("--alpha", type=float, default=1.0)
a = ap.parse_args()

torch.cuda.init()
t0 = time.time()
if a.model == "cnvm":
    params, x0, rate = workloads.headline_cnvm(a.n, alpha=a.alpha)
    model = CNVM(params); kind = "cnvm"; M = 3
    struct, keep = model._struct()
elif a.model == "sis":
    params, x0, rate = workloads.sis_ws(a.n)
    model = CNVM(params); kind = "cnvm"; M = 2
    struct, keep = model._struct()
elif a.model == "complete":
    params, x0, rate = workloads.complete_cnvm(a.n)
    model = CNVM(params); kind = "cnvm"; M = 2
    struct, keep = model._struct()
else:
    params, x0, rate = workloads.ba_cntm(a.n)
    model = CNTM(params); kind = "cntm"; M = 2
    struct = model._struct()
graph = None if (a.model == "complete") else model._graph()
print(f"setup {time.time()-t0:.1f}s  rate={rate:.4g}/unit time")
t_eval = np.linspace(0, a.t_max, a.T)
d_x0 = torch.as_tensor(x0[None, :], device="cuda")
spec = (np.arange(M, dtype=np.int32), 1.0)
kw = dict(cv_spec=spec, want_cv=True, force_global_state=a.global_state,
          **({'warps_per_chain': int(os.environ.get('SP_WPC', '0'))} if kind == 'cnvm' else {}))
tw = time.time()
while time.time() - tw < 1.5:  # an idle B200 needs ~1 s of work to get back to full clocks
    engine.run_native(kind, graph, params.num_agents, struct, d_x0, t_eval, 1, a.runs, 0, a.runs, **kw)
    torch.cuda.synchronize()
for rep in range(a.reps):
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    _, cv = engine.run_native(kind, graph, params.num_agents, struct, d_x0, t_eval, 1234 + rep, a.runs, 0, a.runs,
                              cv_spec=spec, want_cv=True, counters=counters, force_global_state=a.global_state,
                              **({'warps_per_chain': int(os.environ.get('SP_WPC', '0'))} if kind == 'cnvm' else {}))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = counters.cpu().numpy()
    print(f"rep {rep}: {ms:.1f} ms  events={c[0]:.4g}  ev/s={c[0]/ms*1e3:.4g}  acc={c[1]/max(c[0],1):.3f} "
          f"steps={c[2]} extra_rounds/step={c[3]/max(c[2],1):.4f}  mean cv[-1]={cv[0,:,-1].mean(0).cpu().numpy()}")
