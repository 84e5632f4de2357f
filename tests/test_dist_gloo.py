"""CPU, world_size = 2, gloo: the host-side multi-GPU logic (run split, seed broadcast, gather of the
per-rank run slabs, all-reduce of the moment sums).  No kernels are launched here -- the per-rank
compute is replaced by a deterministic stand-in with the same chain-id indexing as the library."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_cv(seed, S, runs, T, D):
    """What a rank would return for global run indices `runs`: a pure function of (seed, state, run)."""
    j = np.arange(S)[:, None, None, None]
    i = np.asarray(runs)[None, :, None, None]
    k = np.arange(T)[None, None, :, None]
    d = np.arange(D)[None, None, None, :]
    return ((seed % 1000) + 1000.0 * j + 10.0 * i + 0.1 * k + 0.01 * d).astype(np.float64)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sponet_b200 import multiprocessing as smp
        from sponet_b200.sample_moments import _allreduce_

        assert smp._dist() == (rank, world)
        # every rank ends up with rank 0's key
        seed = smp._broadcast_seed(0xDEADBEEF00000000 + 17 * (rank + 1))
        assert seed == 0xDEADBEEF00000000 + 17

        S, R, T, D = 2, 7, 3, 2  # 7 runs over 2 ranks -> [4, 3]
        bounds = np.concatenate(([0], np.cumsum(smp._split_runs(R, world))))
        rb, re = int(bounds[rank]), int(bounds[rank + 1])
        local = _fake_cv(seed, S, range(rb, re), T, D)
        full = smp._gather_runs(local, bounds, world)
        np.testing.assert_array_equal(full, _fake_cv(seed, S, range(R), T, D))

        # uint8 state slabs gather too
        xs = (np.arange(S * (re - rb) * T * 5).reshape(S, re - rb, T, 5) % 251).astype(np.uint8) + rank
        allx = smp._gather_runs(xs, bounds, world)
        assert allx.shape == (S, R, T, 5) and allx.dtype == np.uint8
        np.testing.assert_array_equal(allx[:, rb:re], xs)

        # moment partials: integer sums all-reduced in one message
        c = local[0].astype(np.int64)
        s1, s2 = c.sum(0), (c * c).sum(0)
        _allreduce_([s1, s2])
        ref = _fake_cv(seed, S, range(R), T, D)[0].astype(np.int64)
        np.testing.assert_array_equal(s1, ref.sum(0))
        np.testing.assert_array_equal(s2, (ref * ref).sum(0))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
