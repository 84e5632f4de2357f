"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.  Checks that the sharded
ensemble equals the single-GPU ensemble bit for bit and that the all-reduced moments match."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sponet_b200 import CNVMParameters, OpinionShares, sample_many_runs, sample_moments  # noqa: E402
from sponet_b200 import multiprocessing as smp  # noqa: E402
from sponet_b200 import workloads  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    params, x0, _ = workloads.headline_cnvm(3000)
    cv = OpinionShares(3, normalize=True)
    R, T = 37, 6

    # sharded over the ranks, gathered: every rank holds the full result
    _, c_dist = sample_many_runs(params, x0, 2.0, T, R, collective_variable=cv, rng=np.random.default_rng(5))
    _, x_dist = sample_many_runs(params, x0, 1.0, 3, 5, rng=np.random.default_rng(6))
    _, m_dist, v_dist, n_dist = sample_moments(params, x0, 2.0, T, 64, rel_tol=1e9, collective_variable=cv,
                                               rng=np.random.default_rng(7))

    # parameter sweep: points dealt to the ranks, moments + histograms all-reduced on device buffers
    from sponet_b200 import estimate_steady_state, sample_moments_sweep

    plist = [CNVMParameters(num_opinions=3, network=params.network, r=workloads.R3 * (1 + 0.2 * q), r_tilde=workloads.RT3)
             for q in range(3)]
    sw_dist = sample_moments_sweep(plist, x0, 1.0, 4, 21, cv, rng=np.random.default_rng(8), hist_bins=12)
    ss_dist = estimate_steady_state(plist[0], cv, tol_abs=0.02, rng=np.random.default_rng(9), verbose=False, num_chains=16)

    # per-run networks generated on the device: runs sharded, graph c a function of (generator seed, c) only
    from sponet_b200.network_generator import ErdosRenyiGenerator

    def net_params():
        gen = ErdosRenyiGenerator(800, 0.01, rng=np.random.default_rng(21), force_no_isolates=True)
        return CNVMParameters(num_opinions=3, network_generator=gen, r=workloads.R3, r_tilde=workloads.RT3)

    xn = np.random.default_rng(1).integers(0, 3, (2, 800))
    _, n_dist_cv = sample_many_runs(net_params(), xn, 1.5, 4, 11, collective_variable=OpinionShares(3), rng=np.random.default_rng(22))

    # the same calls on ONE GPU (sharding switched off locally)
    real = smp._dist
    smp._dist = lambda: (0, 1)
    import sponet_b200.sample_moments as sm

    sm._dist = smp._dist
    _, c_one = sample_many_runs(params, x0, 2.0, T, R, collective_variable=cv, rng=np.random.default_rng(5))
    _, x_one = sample_many_runs(params, x0, 1.0, 3, 5, rng=np.random.default_rng(6))
    _, m_one, v_one, n_one = sample_moments(params, x0, 2.0, T, 64, rel_tol=1e9, collective_variable=cv,
                                            rng=np.random.default_rng(7))
    sw_one = sample_moments_sweep(plist, x0, 1.0, 4, 21, cv, rng=np.random.default_rng(8), hist_bins=12)
    ss_one = estimate_steady_state(plist[0], cv, tol_abs=0.02, rng=np.random.default_rng(9), verbose=False, num_chains=16)
    _, n_one_cv = sample_many_runs(net_params(), xn, 1.5, 4, 11, collective_variable=OpinionShares(3), rng=np.random.default_rng(22))
    smp._dist = real
    sm._dist = real
    assert n_dist_cv.shape == (2, 11, 4, 3) and np.array_equal(n_dist_cv, n_one_cv), "sharded per-run-network ensemble differs"
    for a, b in zip(sw_dist, sw_one):
        assert np.array_equal(a, b), "sharded sweep differs from the single-GPU sweep"
    assert np.array_equal(ss_dist, ss_one), "sharded steady-state estimate differs"

    assert c_dist.shape == (R, T, 3) and np.array_equal(c_dist, c_one), "sharded CV ensemble differs"
    assert np.array_equal(x_dist, x_one), "sharded state ensemble differs"
    assert n_dist == n_one == 64
    assert np.array_equal(m_dist, m_one) and np.array_equal(v_dist, v_one), "all-reduced moments differ"
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
