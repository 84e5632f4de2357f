import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sponet_b200 import engine, workloads
from sponet_b200.cnvm import CNVM
params, x0, rate = workloads.headline_cnvm(100000)
model = CNVM(params); graph = model._graph(); struct, keep = model._struct()
t_eval = np.linspace(0, 10, 101); spec = (np.arange(3, dtype=np.int32), 1.0)
d_x0 = torch.as_tensor(x0[None, :], device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
counters = torch.zeros(4, dtype=torch.int64, device="cuda")
def step(seed, c=None):
    return engine.run_native("cnvm", graph, 100000, struct, d_x0, t_eval, seed, 10000, 0, 10000, cv_spec=spec, want_cv=True, counters=c)
for w in range(3): step(10+w)
torch.cuda.synchronize()
for mode in ["noflush", "flush", "flush+sleep", "noflush", "smallflush"]:
    ms = []
    for k in range(4):
        if mode.startswith("flush"): flush.fill_(k)
        if mode == "smallflush": flush[:16<<20].fill_(k)
        if mode == "flush+sleep": torch.cuda.synchronize(); time.sleep(0.2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(1000+k, counters); e1.record(); e1.synchronize()
        ms.append(round(e0.elapsed_time(e1), 1))
    print(mode, ms, flush=True)
