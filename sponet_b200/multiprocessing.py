"""Ensemble driver (reference: sponet/multiprocessing.py:18-196).

``sample_many_runs`` keeps the reference's signature and output shapes.  What changes is who does
the work: all ``num_initial_states x num_runs`` realizations go to the GPU in one batched C-ABI
call instead of a process pool of numba loops.

native mode (default)
    Chain (j, i) = (initial state j, run i) draws from the Philox stream keyed by one uint64 taken
    from ``rng`` and countered by ``j * num_runs + i``: results depend on the seed only -- not on
    ``n_jobs``, not on how runs are sharded over GPUs.  In a single process ``n_jobs`` selects how many of
    the visible GPUs share the runs (None / 1: the current one, -1: all), one host thread per device.  An unweighted
    ``OpinionShares`` CV is reduced inside the simulation kernel (complete graphs take the count-only
    kernel); other CVs are evaluated on the GPU from snapshot slabs, run-chunk by run-chunk.
    Under ``torch.distributed`` (one process per GPU) the runs are split over the ranks like
    ``_split_runs`` and the pieces are all-gathered, so every rank returns the full array.

replay mode
    Reproduces the reference bit for bit, including its stream discipline: ``n_jobs in (None, 1)``
    threads ONE generator through all runs (multiprocessing.py:171-178), otherwise ``rng.spawn(n_jobs)``
    gives one stream per chunk (:90-103).  One GPU warp replays each stream.
"""
from __future__ import annotations

import os
import weakref

import numpy as np
from numpy.random import Generator, default_rng
from numpy.typing import ArrayLike, NDArray

from . import engine
from .cntm.model import CNTM
from .cntm.parameters import CNTMParameters
from .cnvm.model import CNVM
from .cnvm.parameters import CNVMParameters
from .collective_variables import CollectiveVariable, OpinionShares
from .parameters import Parameters
from .utils import t_eval_to_ndarray

# device memory budget for snapshot slabs when a CV has to be evaluated from stored states
_SLAB_BYTES = int(os.environ.get("SPONET_B200_SLAB_BYTES", 2 << 30))
# Complete graphs with a counts-only CV: the thread-per-chain count kernel needs tens of thousands of
# chains to fill a B200; below that the warp-per-chain agent-level kernel is faster.  The choice
# depends on the TOTAL number of chains of the call so that it does not change with the GPU count.
_COUNT_ONLY_MIN_CHAINS = 20_000


def use_count_only(model, kind, spec, total_chains: int) -> bool:
    p = model.params
    return (kind == "cnvm" and model.is_complete and p.num_opinions <= 16 and spec is not None and not spec.grouped
            and spec.divisors is None and total_chains >= _COUNT_ONLY_MIN_CHAINS)


def _split_runs(num_runs: int, num_chunks: int) -> np.ndarray:
    """Near-equal chunk sizes, larger chunks first (reference :190-196): (20, 6) -> [4,4,3,3,3,3]."""
    base, extra = divmod(num_runs, num_chunks)
    chunks = np.full(num_chunks, base, dtype=int)
    chunks[:extra] += 1
    return chunks


# The reference rebuilds the model in every worker (multiprocessing.py:146-151).  Here that would mean
# re-deriving the CSR and re-uploading it to HBM on every call, so models are kept per parameter
# object for as long as the network / alpha they were built from are unchanged.
_MODEL_CACHE: dict[int, tuple] = {}


def _network_digest(net) -> int:
    """Cheap content hash of the adjacency (a network rewired IN PLACE keeps its id): CRC of the degree
    sequence and of the concatenated neighbour ids."""
    import zlib

    if net is None:
        return 0
    if hasattr(net, "adj"):  # networkx graph (CNTM parameters)
        deg = np.fromiter((d for _, d in net.degree()), dtype=np.int64)
        nbr = np.fromiter((v for _, nb in net.adjacency() for v in nb), dtype=np.int64, count=int(deg.sum()))
    else:
        deg = np.fromiter(map(len, net), dtype=np.int64, count=len(net))
        try:  # rows that are contiguous arrays: one C-level pass over their buffers (~10 ms per 1e5 rows instead of ~80)
            return zlib.crc32(b"".join(net), zlib.crc32(deg.tobytes()))
        except (TypeError, BufferError, ValueError):
            nbr = np.concatenate([np.asarray(nb, dtype=np.int64) for nb in net]) if len(net) else np.zeros(0, np.int64)
    return zlib.crc32(nbr.tobytes(), zlib.crc32(deg.tobytes()))


def _fingerprint(params):
    net = getattr(params, "network", None)
    if isinstance(params, CNVMParameters):
        return (id(net), _network_digest(net), params.alpha, params.num_agents)
    return (id(net), _network_digest(net), params.r, params.r_tilde)


def _make_model(params: Parameters):
    if isinstance(params, CNVMParameters):
        kind, cls = "cnvm", CNVM
    elif isinstance(params, CNTMParameters):
        kind, cls = "cntm", CNTM
    else:
        raise ValueError("Parameters not valid.")
    if getattr(params, "network_generator", None) is not None:
        return cls(params), kind
    key, fp = id(params), _fingerprint(params)
    hit = _MODEL_CACHE.get(key)
    if hit is not None and hit[0]() is params and hit[2] == fp:
        return hit[1], kind
    model = cls(params)
    if len(_MODEL_CACHE) >= 8:
        _MODEL_CACHE.pop(next(iter(_MODEL_CACHE)))
    _MODEL_CACHE[key] = (weakref.ref(params), model, fp)
    return model, kind


def _dist():
    """(rank, world) of an initialised torch.distributed job, else (0, 1)."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def sample_many_runs(
    params: Parameters,
    initial_states: ArrayLike,
    t_max: float,
    t_eval: ArrayLike,
    num_runs: int,
    n_jobs: int | None = None,
    collective_variable: CollectiveVariable | None = None,
    rng: Generator = default_rng(),
    progress_bar: bool = False,
    mode: str | None = None,
) -> tuple[NDArray, NDArray]:
    """Sample ``num_runs`` realizations per initial state.

    Returns ``(t_out, x_out)`` with ``t_out.shape == (T,)`` and ``x_out.shape == (S, R, T, N)`` (dtype of
    the opinions) or ``(S, R, T, D)`` float64 with a collective variable; the leading axis is dropped
    for a single 1-D initial state.
    """
    on_device = hasattr(initial_states, "is_cuda") and initial_states.is_cuda
    if on_device:  # states sampled on the GPU (sponet_b200.states, device=True): no host copy
        return _sample_device_states(params, initial_states, t_max, t_eval, num_runs, collective_variable, rng, progress_bar, mode)
    initial_states = np.array(initial_states, ndmin=1)
    is_1d = initial_states.ndim == 1
    if is_1d:
        initial_states = initial_states[None, :]
    t_out = t_eval_to_ndarray(t_eval, t_max)
    model, kind = _make_model(params)
    mode = engine.resolve_mode(mode)
    dtype = np.min_scalar_type(params.num_opinions - 1)
    _check_states(initial_states, params.num_opinions)

    per_run_network = getattr(params, "network_generator", None) is not None
    if mode == "replay" or params.num_opinions > 256:
        x_out = _sample_replay(model, kind, initial_states.astype(dtype), t_max, t_out, num_runs, n_jobs,
                               collective_variable, rng, progress_bar, per_run_network)
    elif per_run_network:
        x_out = _sample_native_per_run_network(model, kind, initial_states.astype(np.uint8), t_out, num_runs,
                                               collective_variable, rng, progress_bar).astype(
            dtype if collective_variable is None else np.float64, copy=False)
    else:
        x_out = _sample_native(model, kind, initial_states.astype(np.uint8), t_out, num_runs,
                               collective_variable, rng, progress_bar, n_jobs).astype(
            dtype if collective_variable is None else np.float64, copy=False)
    return t_out, (x_out[0] if is_1d else x_out)


def _sample_device_states(params, d_states, t_max, t_eval, num_runs, cv, rng, progress_bar, mode):
    """sample_many_runs for initial states that already live on the GPU (CUDA uint8 tensor [S, N] or [N]): native
    mode with an in-kernel collective variable only -- the case the device-side samplers exist for."""
    import torch

    model, kind = _make_model(params)
    if engine.resolve_mode(mode) != "native" or getattr(params, "network_generator", None) is not None:
        raise ValueError("device-resident initial states need native mode and a fixed network")
    spec = _kernel_cv_spec(cv, params.num_agents) if cv is not None else None
    if spec is not None and (spec.grouped or spec.divisors is not None) and kind != "cnvm":
        spec = None
    if spec is None:
        raise ValueError("device-resident initial states need a collective variable the kernels evaluate")
    if d_states.dtype != torch.uint8:
        raise ValueError("device-resident initial states must be uint8")
    is_1d = d_states.ndim == 1
    st = (d_states[None, :] if is_1d else d_states).contiguous()
    if int(st.max().item()) >= params.num_opinions:
        raise ValueError(f"initial states must hold opinions in [0, {params.num_opinions})")
    t_out = t_eval_to_ndarray(t_eval, t_max)
    seed = engine.draw_seed(rng)
    rank, world = _dist()
    if world > 1:
        seed = _broadcast_seed(seed)
    bounds = np.concatenate(([0], np.cumsum(_split_runs(num_runs, world))))
    rb, re = int(bounds[rank]), int(bounds[rank + 1])
    _, local = _native_call(model, kind, st, t_out, seed, num_runs, rb, re, cv_spec=spec, want_cv=True)
    x_out = _gather_runs(local, bounds, world) if world > 1 else local.cpu().numpy()
    return t_out, (x_out[0] if is_1d else x_out)


def _check_states(states: np.ndarray, num_opinions: int) -> None:
    """Opinions outside [0, M) would be bit-packed into a neighbouring agent's field; the reference indexes
    prob_imit[x, .] with them unchecked (cnvm/model.py:288).  Rejected here, before the uint8 cast."""
    if states.size and (states.min() < 0 or states.max() >= num_opinions):
        raise ValueError(f"initial states must hold opinions in [0, {num_opinions})")


# ---------------------------------------------------------------------------------------------
# native
# ---------------------------------------------------------------------------------------------
def _kernel_cv_spec(cv, num_agents):
    """KernelCV if the simulation kernels can evaluate `cv` on the fly (OpinionShares family with
    integer weights, by-degree groups, composites over one table), else None."""
    fn = getattr(cv, "_kernel_spec", None)
    return fn(num_agents) if fn is not None else None


def _native_call(model, kind, x_init, t_out, seed, R_total, rb, re, **kw):
    p = model.params
    if kind == "cnvm":
        struct, keep = model._struct()
        out = engine.run_native("cnvm", model._graph(), p.num_agents, struct, x_init, t_out, seed, R_total, rb, re, **kw)
        del keep
        return out
    return engine.run_native("cntm", model._graph(), p.num_agents, model._struct(), x_init, t_out, seed, R_total, rb, re, **kw)


def _local_devices(n_jobs) -> int:
    """GPUs a SINGLE process spreads an ensemble over: ``n_jobs`` keeps its meaning of "how many workers" --
    None / 1: the current device only (like the reference's serial path), -1: every visible device, k: k devices.
    (Under torch.distributed every rank owns one device and ``n_jobs`` is ignored.)"""
    if n_jobs is None or n_jobs == 1:
        return 1
    try:
        import torch

        have = torch.cuda.device_count()
    except Exception:
        return 1
    return have if n_jobs == -1 else max(1, min(int(n_jobs), have))


def _multi_device_runs(model, kind, states, t_out, seed, num_runs, rb, re, spec, num_devices):
    """Runs [rb, re) split like ``_split_runs`` over ``num_devices`` GPUs of this process: one host thread per
    device (the C-ABI call releases the GIL and is re-entrant per device; the graph is uploaded once per device),
    host buffers in and out, slabs concatenated in run order -- chain ids do not depend on the split, so the result
    equals the single-device one bit for bit."""
    import threading

    import torch

    b = rb + np.concatenate(([0], np.cumsum(_split_runs(re - rb, num_devices))))
    parts, errors = [None] * num_devices, []
    cur = torch.cuda.current_device()

    def work(d):
        try:
            torch.cuda.set_device(d)
            _, parts[d] = _native_call(model, kind, states, t_out, seed, num_runs, int(b[d]), int(b[d + 1]), cv_spec=spec, want_cv=True)
        except Exception as e:  # re-raised in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d,)) for d in range(num_devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.set_device(cur)
    if errors:
        raise errors[0]
    return np.concatenate(parts, axis=1)


def _sample_native(model, kind, states, t_out, num_runs, cv, rng, progress_bar, n_jobs=None):
    p = model.params
    S, N, T = states.shape[0], p.num_agents, t_out.size
    seed = engine.draw_seed(rng)
    rank, world = _dist()
    if world > 1:  # every rank must use the same key: take rank 0's
        seed = _broadcast_seed(seed)
    bounds = np.concatenate(([0], np.cumsum(_split_runs(num_runs, world))))
    rb, re = int(bounds[rank]), int(bounds[rank + 1])
    states = np.ascontiguousarray(states)
    pbar = _progress(progress_bar, S * (re - rb))
    k_states = states  # what the kernel call gets: under NCCL a device copy, so that its output stays on the GPU
    if world > 1 and _nccl():
        import torch

        k_states = torch.as_tensor(states, device="cuda")

    spec = _kernel_cv_spec(cv, N)
    if spec is not None and (spec.grouped or spec.divisors is not None) and kind != "cnvm":
        spec = None  # grouped / weighted tables are kept by the CNVM agent-level kernel only
    if spec is not None and spec.edge and p.num_opinions != 2:
        raise ValueError("Interfaces / Propensities can only be used for 2 opinions.")
    if spec is not None:
        if not spec.edge and use_count_only(model, kind, spec, S * num_runs):
            counts0 = np.stack([np.bincount(s, minlength=p.num_opinions) for s in states]).astype(np.int64)
            struct, keep = model._struct()
            _, local = engine.run_native("cnvm_counts", None, N, struct, None, t_out, seed, num_runs, rb, re,
                                         cv_spec=spec, want_cv=True, counts_init=counts0)
            del keep
        elif world == 1 and _local_devices(n_jobs) > 1 and re - rb >= 2:
            local = _multi_device_runs(model, kind, states, t_out, seed, num_runs, rb, re, spec, _local_devices(n_jobs))
        else:
            _, local = _native_call(model, kind, k_states, t_out, seed, num_runs, rb, re, cv_spec=spec, want_cv=True)
        _tick(pbar, S * (re - rb))
    elif cv is None:
        # full trajectories: produced in run chunks so that the device-side staging stays bounded
        local = np.empty((S, re - rb, T, N), dtype=np.uint8)
        chunk = max(1, int(_SLAB_BYTES // max(1, S * T * N)))
        for lo in range(rb, re, chunk):
            hi = min(re, lo + chunk)
            part, _ = _native_call(model, kind, states, t_out, seed, num_runs, lo, hi, want_x=True)
            local[:, lo - rb : hi - rb] = part
            _tick(pbar, S * (hi - lo))
    else:
        local = _native_generic_cv(model, kind, states, t_out, seed, num_runs, rb, re, cv, pbar)
    _close(pbar)
    return _gather_runs(local, bounds, world) if world > 1 else local


def _sample_native_per_run_network(model, kind, states, t_out, num_runs, cv, rng, progress_bar):
    """A NetworkGenerator in the parameters: a new network for every (initial state, run) pair, sampled in the
    reference's order (states outer, runs inner: multiprocessing.py:173-178 with cnvm/model.py:95-96).  Every
    pair is one native chain on its own graph; chain id = j * num_runs + i keys its Philox stream, so the
    result does not depend on how the pairs are spread over ranks.  Each rank draws the networks of all pairs
    from its own copy of the generator (as every worker process of the reference does with its pickled copy)
    and simulates its share."""
    p = model.params
    S, N, T = states.shape[0], p.num_agents, t_out.size
    seed = engine.draw_seed(rng)
    rank, world = _dist()
    if world > 1:
        seed = _broadcast_seed(seed)
    pairs = S * num_runs
    spec = _kernel_cv_spec(cv, N) if cv is not None else None
    if spec is not None and (spec.edge and p.num_opinions != 2):
        raise ValueError("Interfaces / Propensities can only be used for 2 opinions.")
    if spec is not None and spec.edge:
        spec = None  # an edge CV carries its own (fixed) graph; evaluated from the returned states below
    gen = p.network_generator
    if kind == "cnvm" and hasattr(gen, "device_batch") and not _FORCE_HOST_NETWORKS:
        # the generator makes its graphs on the GPU: the whole ensemble in a few launches, one graph per chain
        return _per_run_network_batched(model, states, t_out, seed, num_runs, cv, spec, progress_bar)
    bounds = np.concatenate(([0], np.cumsum(_split_runs(pairs, world))))
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    D = N if cv is None else cv.dimension
    flat = np.zeros((pairs, T, D), dtype=np.uint8 if cv is None else np.float64)
    pbar = _progress(progress_bar, hi - lo)
    if spec is not None:
        flat[lo:hi] = _per_run_network_pipeline(model, kind, states, t_out, seed, pairs, lo, hi, cv, pbar)
    else:
        for c in range(pairs):
            model.update_network()  # every rank walks the same generator sequence
            if not (lo <= c < hi):
                continue
            j = c // num_runs
            x0 = np.ascontiguousarray(states[j : j + 1])
            out, _ = _native_call(model, kind, x0, t_out, seed, pairs, c, c + 1, want_x=True)
            flat[c] = out[0, 0] if cv is None else np.asarray(cv(out[0, 0]))
            _tick(pbar, 1)
    _close(pbar)
    if world > 1:
        flat = _gather_runs(flat[None, lo:hi], bounds, world)[0]
    return flat.reshape(S, num_runs, T, D)


_NETWORK_STREAMS = 8
_FORCE_HOST_NETWORKS = False  # tests: take the host-generated pipeline although the generator has device_batch
_BATCH_BYTES = None           # device memory one batch of generated graphs may take (None: from the device, see below)
_POOLED = [0]                 # bytes of the last batch: kept by the library's pool, so "free" understates what is available


def _pooled_bytes_hint() -> int:
    return _POOLED[0]


def _per_run_network_batched(model, states, t_out, seed, num_runs, cv, spec, progress_bar):
    """Per-run networks from a generator with ``device_batch`` (the generators of sponet_b200.network_generator): the
    graphs of a block of (initial state, run) pairs are generated on the GPU and the block's chains run in ONE launch,
    chain (j, i) on graph index j * num_runs + i.  Graph and Philox stream of a pair depend on the two seeds and on the
    pair's index only: the result is the same for every block size and number of ranks.  Runs are split over ranks
    like ``_split_runs`` (multiprocessing.py:190-196)."""
    p = model.params
    gen = p.network_generator
    S, N, T = states.shape[0], p.num_agents, t_out.size
    gseed = gen.draw_batch_seed()
    rank, world = _dist()
    if world > 1:
        gseed = _broadcast_seed(gseed)
    bounds = np.concatenate(([0], np.cumsum(_split_runs(num_runs, world))))
    rb, re = int(bounds[rank]), int(bounds[rank + 1])
    D = N if cv is None else cv.dimension
    local = np.empty((S, re - rb, T, D), dtype=np.uint8 if cv is None else np.float64)
    want_x = cv is None or spec is None  # a CV the kernels cannot evaluate: from the returned states, pair by pair
    # graphs per block: device memory (rows 8 B + record 64 B + 12 B of build scratch per vertex, 4 B per entry), the
    # 32-bit row / entry indices of a batch, and -- for full trajectories -- the output slab
    mean_entries = N * (gen.expected_mean_degree() + 2.0)
    per_graph = (84.0 + getattr(gen, "build_bytes_per_vertex", 0.0)) * N + 4.0 * mean_entries + (T * N if want_x else 0)
    if p.alpha != 1:
        per_graph += (8.0 + 16.0 - 64.0) * N  # alias table + its construction scratch instead of the records
    budget = _BATCH_BYTES
    if budget is None:  # a third of the device (60 GB of a B200's 180), never more than is free right now
        import torch

        free, total = torch.cuda.mem_get_info()
        budget = min(total // 3, int(0.6 * free) + _pooled_bytes_hint())
    max_graphs = int(min(budget // per_graph, (2**32 - 1) // (16 * N), (2**32 - 1) // (1.2 * mean_entries)))
    max_graphs = max(1, max_graphs)
    s_blk = min(S, max_graphs)
    r_blk = max(1, max_graphs // s_blk)
    r_blk = max(1, -(-(re - rb) // max(1, -(-(re - rb) // r_blk))))  # equal blocks instead of full ones and a remainder
    struct, keep = model._struct()
    pbar = _progress(progress_bar, S * (re - rb))
    for s0 in range(0, S, s_blk):
        s1 = min(S, s0 + s_blk)
        st = np.ascontiguousarray(states[s0:s1])
        for lo in range(rb, re, r_blk):
            hi = min(re, lo + r_blk)
            # (state block offset folded into the run offset: index = (s0 + j) * num_runs + i for the block's j)
            batch = gen.device_batch(gseed, s1 - s0, num_runs, s0 * num_runs + lo, s0 * num_runs + hi, records=p.alpha == 1)
            try:
                if p.alpha != 1:  # agent weights degree ** (1 - alpha), event rate and noise probability per graph
                    batch.set_weights(p.alpha)
                out_x, out_cv = engine.run_native("cnvm", batch, N, struct, st, t_out, seed, num_runs, lo, hi,
                                                  cv_spec=None if want_x else spec, want_cv=not want_x, want_x=want_x,
                                                  chain_offset=s0 * num_runs)
            finally:
                batch.close()
            _POOLED[0] = int(per_graph * (s1 - s0) * (hi - lo))
            if cv is None:
                local[s0:s1, lo - rb : hi - rb] = out_x
            elif want_x:
                for j in range(s1 - s0):
                    for i in range(hi - lo):
                        local[s0 + j, lo - rb + i] = np.asarray(cv(out_x[j, i]))
            else:
                local[s0:s1, lo - rb : hi - rb] = out_cv
            _tick(pbar, (s1 - s0) * (hi - lo))
    del keep
    _close(pbar)
    return _gather_runs(local, bounds, world) if world > 1 else local


def _per_run_network_pipeline(model, kind, states, t_out, seed, pairs, lo, hi, cv, pbar):
    """Chains c in [lo, hi), each on the c-th network of the generator, with the CV reduced in-kernel.

    One chain is one CTA, so a launch per chain would use 1 of 148 SMs: the chains go round-robin onto
    ``_NETWORK_STREAMS`` CUDA streams with device-resident inputs and outputs -- the C-ABI call returns as soon
    as the kernel is queued -- so up to that many graphs are simulated concurrently while the host generates
    (networkx) and uploads the next one.  A graph handle is released when its stream has drained."""
    import torch

    p = model.params
    N, T = p.num_agents, t_out.size
    num_runs = pairs // states.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    streams = [torch.cuda.Stream(device=dev) for _ in range(_NETWORK_STREAMS)]
    alive = [[] for _ in streams]  # graph handles (and the buffers their launch reads) per stream
    d_states = torch.as_tensor(np.ascontiguousarray(states), device=dev)
    out_d = torch.zeros((max(hi - lo, 1), T, cv.dimension), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for c in range(pairs):
        model.update_network()  # every rank walks the same generator sequence
        if not (lo <= c < hi):
            continue
        k = (c - lo) % len(streams)
        st = streams[k]
        if alive[k]:  # the previous chain of this stream: wait for it, then free its graph
            st.synchronize()
            for h in alive[k][0]:
                h.close()
            alive[k] = []
        spec = _kernel_cv_spec(cv, N)  # (per network: degree-weighted CVs depend on it)
        graph = model._graph()
        struct, keep = model._struct() if kind == "cnvm" else (model._struct(), None)
        j = c // num_runs
        with torch.cuda.stream(st):
            _, res = engine.run_native(kind, graph, N, struct, d_states[j : j + 1], t_out, seed, pairs, c, c + 1,
                                       cv_spec=spec, want_cv=True)
            out_d[c - lo].copy_(res[0, 0], non_blocking=True)
        alive[k] = [model._graphs.detach(), keep, res, spec]
        _tick(pbar, 1)
    torch.cuda.synchronize()
    for a in alive:
        if a:
            for h in a[0]:
                h.close()
    return out_d[: hi - lo].cpu().numpy()


def _native_generic_cv(model, kind, states, t_out, seed, num_runs, rb, re, cv, pbar):
    """CV evaluated from snapshot slabs that stay on the GPU; runs are processed in memory-sized chunks."""
    import torch

    p = model.params
    S, N, T = states.shape[0], p.num_agents, t_out.size
    out = np.empty((S, re - rb, T, cv.dimension), dtype=np.float64)
    chunk = max(1, int(_SLAB_BYTES // max(1, S * T * N)))
    d_states = torch.as_tensor(states, device="cuda")
    for lo in range(rb, re, chunk):
        hi = min(re, lo + chunk)
        slab, _ = _native_call(model, kind, d_states, t_out, seed, num_runs, lo, hi, want_x=True)
        try:
            c = cv(slab.reshape(-1, N))  # GPU path of the OpinionShares family
        except (TypeError, AttributeError):  # user-defined CV: host evaluation on the copied slab
            c = cv(slab.reshape(-1, N).cpu().numpy())
        c = c.cpu().numpy() if hasattr(c, "cpu") else np.asarray(c)
        out[:, lo - rb : hi - rb] = c.reshape(S, hi - lo, T, cv.dimension)
        _tick(pbar, S * (hi - lo))
    return out


def _nccl() -> bool:
    try:
        import torch.distributed as dist

        return dist.get_backend() == "nccl"
    except Exception:
        return False


def _broadcast_seed(seed: int) -> int:
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, seed >> 63], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=0)
    return int(t[0].item()) | (int(t[1].item()) << 63)


def _gather_runs(local, bounds: np.ndarray, world: int) -> np.ndarray:
    """All-gather the per-rank run slices (axis 1) into the full array on every rank.  ``local`` is this rank's
    slab, a numpy array or -- on the native path under NCCL -- the CUDA tensor the kernel wrote: it goes into the
    all-gather as it is and the full array is copied to the host once."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    sizes = np.diff(bounds)
    width = int(sizes.max())
    loc = local if hasattr(local, "data_ptr") else torch.from_numpy(np.ascontiguousarray(local))
    shape = list(loc.shape)
    shape[1] = width
    if loc.shape[1] == width and str(loc.device).startswith(dev):
        pad = loc.contiguous()
    else:
        pad = torch.zeros(shape, dtype=loc.dtype, device=dev)
        pad[:, : loc.shape[1]] = loc.to(dev)
    flat = torch.empty([world * shape[0]] + shape[1:], dtype=loc.dtype, device=dev)  # concatenated along axis 0
    dist.all_gather_into_tensor(flat, pad)
    full = flat.view([world] + shape)
    if all(int(sz) == width for sz in sizes):  # equal shares: [world, S, w, ...] -> [S, world * w, ...] in one copy
        return full.movedim(0, 1).reshape([shape[0], world * width] + shape[2:]).cpu().numpy()
    host = full.cpu().numpy()
    return np.concatenate([host[r][:, : int(sizes[r])] for r in range(world)], axis=1)


# ---------------------------------------------------------------------------------------------
# replay
# ---------------------------------------------------------------------------------------------
def _sample_replay(model, kind, states, t_max, t_out, num_runs, n_jobs, cv, rng, progress_bar, per_run_network):
    p = model.params
    S, N, T = states.shape[0], p.num_agents, t_out.size
    D = N if cv is None else cv.dimension
    x_out = np.zeros((S, num_runs, T, D), dtype=states.dtype if cv is None else np.float64)

    if per_run_network:
        # A NetworkGenerator resamples the graph inside every simulate() call (cnvm/model.py:95-96):
        # run the realizations one by one on a single generator, exactly like _Worker.__call__.
        pbar = _progress(progress_bar, S * num_runs)
        for j in range(S):
            for i in range(num_runs):
                _, x = model.simulate(t_max, x_init=states[j], t_eval=t_out, rng=rng, mode="replay")
                x_out[j, i] = x if cv is None else cv(x)
                _tick(pbar, 1)
        _close(pbar)
        return x_out

    if n_jobs is None or n_jobs == 1:
        rngs, work = [rng], [(0, S, 0, num_runs)]
    else:
        if n_jobs == -1:
            n_jobs = os.cpu_count()
            if n_jobs is None:
                raise RuntimeError("Could not determine number of available CPUs.")
        rngs = rng.spawn(n_jobs)
        if num_runs >= S:  # split along runs (reference :96-99)
            b = np.concatenate(([0], np.cumsum(_split_runs(num_runs, n_jobs))))
            work = [(0, S, int(b[w]), int(b[w + 1])) for w in range(n_jobs)]
        else:  # split along initial states (reference :100-103, np.array_split sizes)
            b = np.concatenate(([0], np.cumsum([len(c) for c in np.array_split(np.arange(S), n_jobs)])))
            work = [(int(b[w]), int(b[w + 1]), 0, num_runs) for w in range(n_jobs)]

    if kind == "cnvm":
        struct, keep = model._struct()
        rate, _ = model.event_rates()
    else:
        struct, keep, rate = model._struct(), None, model.next_event_rate
    budget = engine.stream_budget((t_out[-1] - t_out[0]) / rate + 1, model._draws_per_event())
    _, xs, _ = engine.run_replay(kind, model._graph(), N, struct, np.ascontiguousarray(states), num_runs, t_out,
                                 rngs, work, budget)
    del keep
    if cv is None:
        return xs
    for j in range(S):
        for i in range(num_runs):
            x_out[j, i] = cv(xs[j, i])
    return x_out


# ---------------------------------------------------------------------------------------------
# progress bar plumbing (reference shows tqdm for worker 0 only, :92-94, :169, :184-185)
# ---------------------------------------------------------------------------------------------
def _progress(enabled: bool, total: int):
    if not enabled:
        return None
    from tqdm import tqdm

    return tqdm(total=total)


def _tick(pbar, n):
    if pbar is not None:
        pbar.update(n)


def _close(pbar):
    if pbar is not None:
        pbar.close()
