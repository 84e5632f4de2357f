cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_graphgen.py -x -q 2>&1 | tail -3
timeout 600 python - <<'PY'
import time, numpy as np, torch
from sponet_b200 import engine, CNVMParameters, RandomRegularGenerator, OpinionShares, sample_many_runs
N=100000
for G in (148, 1184):
    torch.cuda.synchronize(); t0=time.perf_counter()
    b=engine.GraphBatch.random_regular(N,10,3,1,G,0,G); torch.cuda.synchronize()
    dt=time.perf_counter()-t0; info=b.info(); print(f"regular generate {G} graphs: {dt*1e3:.1f} ms = {dt/G*1e6:.0f} us per graph; degrees {info['min_degree']}..{info['max_degree']}", flush=True); b.close()
kw=dict(num_opinions=3, r=[[0,1,2],[1,0,1],[2,0,0]], r_tilde=[[0,.2,.1],[0,0,.1],[.1,.2,0]])
x0=np.random.default_rng(0).integers(0,3,N)
R=2368
for rep in range(2):
    p=CNVMParameters(network_generator=RandomRegularGenerator(N,10,rng=np.random.default_rng(1)), **kw)
    t0=time.perf_counter(); _,c=sample_many_runs(p,x0,10.0,11,R,collective_variable=OpinionShares(3),rng=np.random.default_rng(2)); dt=time.perf_counter()-t0
    print(f"regular per-run networks: R={R} t_max=10: {dt:.3f} s, {R*2.6e6/dt:.3e} events/s", flush=True)
PY
