cd $GRAFT_REPO_ROOT
export PYTHONPATH=$PWD
timeout 800 python -m pytest tests/test_gpu_statistics.py -x -q -s -k "per_run" 2>&1 | grep -v "^$" | tail -7
timeout 600 python - <<'PY'
import time, numpy as np
from sponet_b200 import CNVMParameters, ErdosRenyiGenerator, OpinionShares, sample_many_runs
N=100000; R=1184
kw=dict(num_opinions=3, r=[[0,1,2],[1,0,1],[2,0,0]], r_tilde=[[0,.2,.1],[0,0,.1],[.1,.2,0]])
x0=np.random.default_rng(0).integers(0,3,N)
for alpha in (1.0, 0.5):
    for rep in range(2):
        p=CNVMParameters(network_generator=ErdosRenyiGenerator(N,10/(N-1),rng=np.random.default_rng(1),force_no_isolates=True), alpha=alpha, **kw)
        t0=time.perf_counter(); _,c=sample_many_runs(p,x0,1.0,11,R,collective_variable=OpinionShares(3),rng=np.random.default_rng(2)); dt=time.perf_counter()-t0
    print(f"alpha={alpha}: ER per-run networks R={R} t_max=1 (units of the model's clock): {dt:.3f} s", flush=True)
PY
