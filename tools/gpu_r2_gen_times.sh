cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_graphgen.py -x -q 2>&1 | tail -2
timeout 600 python - <<'PY'
import time, torch
from sponet_b200 import engine
N=100000
mk={"er": lambda G: engine.GraphBatch(N,1e-4,3,1,G,0,G,force_no_isolates=True), "ba": lambda G: engine.GraphBatch.barabasi_albert(N,5,3,1,G,0,G),
    "ws": lambda G: engine.GraphBatch.watts_strogatz(N,10,0.1,3,1,G,0,G), "regular": lambda G: engine.GraphBatch.random_regular(N,10,3,1,G,0,G)}
for name,f in mk.items():
    for G in (148,1184):
        for rep in range(2):
            torch.cuda.synchronize(); t0=time.perf_counter(); b=f(G); torch.cuda.synchronize(); dt=time.perf_counter()-t0; b.close()
        print(f"{name}: {G} graphs of 1e5 vertices in {dt*1e3:.1f} ms = {dt/G*1e6:.0f} us per graph", flush=True)
PY
