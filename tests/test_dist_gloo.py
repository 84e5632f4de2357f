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
        # parameter sweep: the (point, run) range dealt to the ranks in at most three pieces each; the integer
        # partials of all ranks meet in one all-reduce of (CPU stand-ins of) the device buffers
        import torch

        from sponet_b200.sample_moments import _allreduce_device_, sweep_pieces

        P, Rs = 5, 9  # 45 work items over 2 ranks: 23 + 22 -> rank 0 ends inside point 2
        pieces = sweep_pieces(P, Rs, rank, world)
        assert len(pieces) <= 3
        sums = torch.zeros((2, P, 4, 3), dtype=torch.int64)
        hist = torch.zeros((P, 4, 3, 8), dtype=torch.int64)
        for (pb, pe, r0, r1) in pieces:
            for pt in range(pb, pe):
                for run in range(r0, r1):
                    v = 100 * pt + run  # stand-in of a chain's CV value
                    sums[0, pt] += v
                    sums[1, pt] += v * v
                    hist[pt, :, :, v % 8] += 1
        _allreduce_device_([sums, hist])
        for pt in range(P):
            vals = np.array([100 * pt + run for run in range(Rs)])
            assert int(sums[0, pt, 0, 0]) == vals.sum() and int(sums[1, pt, 3, 2]) == (vals * vals).sum()
            np.testing.assert_array_equal(hist[pt, 0, 0].numpy(), np.bincount(vals % 8, minlength=8))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sweep_pieces_cover_every_item_once():
    from sponet_b200.sample_moments import sweep_pieces

    for P, R, world in [(256, 100000, 8), (5, 9, 2), (3, 10, 8), (1, 7, 3), (16, 4, 5), (2, 1, 4)]:
        seen = np.zeros((P, R), dtype=int)
        for rank in range(world):
            pieces = sweep_pieces(P, R, rank, world)
            assert len(pieces) <= 3
            for (pb, pe, r0, r1) in pieces:
                assert 0 <= pb < pe <= P and 0 <= r0 < r1 <= R and (pe - pb == 1 or (r0, r1) == (0, R))
                seen[pb:pe, r0:r1] += 1
        assert (seen == 1).all(), (P, R, world)
    # 256 points on 8 ranks: whole points only, 32 each
    assert sweep_pieces(256, 100000, 3, 8) == [(96, 128, 0, 100000)]
    # cost-weighted cuts: every (point, run) still exactly once, and the ranks' costs are balanced
    rng = np.random.default_rng(0)
    for P, R, world in [(256, 1000, 8), (7, 50, 3), (2, 9, 5)]:
        w = rng.uniform(0.5, 2.0, P)
        seen = np.zeros((P, R), dtype=int)
        cost = np.zeros(world)
        for rank in range(world):
            for (pb, pe, r0, r1) in sweep_pieces(P, R, rank, world, w):
                seen[pb:pe, r0:r1] += 1
                cost[rank] += w[pb:pe].sum() * (r1 - r0)
        assert (seen == 1).all(), (P, R, world)
        assert cost.max() - cost.min() <= 2.1 * w.max(), (cost, P, R, world)  # within a run or two of each other


def test_two_rank_gloo_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
