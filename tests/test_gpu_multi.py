"""GPU, >= 2 devices: runs sharded over ranks (one process per GPU, NCCL) give the single-GPU result
bit for bit, and sample_moments' all-reduced sums match.  Skipped on a single-GPU box; the host-side
logic is covered on CPU by tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_sharding_matches_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(here, "_dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_single_process_n_jobs_spreads_runs_over_devices():
    """n_jobs = -1 in ONE process: runs split over all visible GPUs (a host thread per device), same result."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from sponet_b200 import OpinionShares, sample_many_runs, workloads

    params, x0, _ = workloads.headline_cnvm(3000)
    cv = OpinionShares(3)
    _, one = sample_many_runs(params, x0, 1.0, 5, 41, collective_variable=cv, rng=np.random.default_rng(5))
    _, many = sample_many_runs(params, x0, 1.0, 5, 41, n_jobs=-1, collective_variable=cv, rng=np.random.default_rng(5))
    np.testing.assert_array_equal(one, many)
