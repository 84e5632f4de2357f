#!/usr/bin/env python
"""Headline benchmark: agent-update events/s of the CNVM ensemble on an Erdos-Renyi N=1e5 network.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric, SURVEY.md section 8(d) "H", BASELINE.md section 3 headline row):
CNVM, M = 3, ER G(1e5, 10/(N-1)) seed 12345 (nnz 998 782), fixture rates (r_imit 2, r_noise 0.6),
alpha = 1, x_init = default_rng(0).integers(0, 3, N), t_max = 100, 101 output times, OpinionShares(3)
reduced in-kernel, 10^4 realizations per GPU.  One "step" = one pass of the hot path over that batch
(~2.6e10 candidate events per GPU).  An "event" is one iteration of the reference while-loop
(sponet/cnvm/model.py:274-306): one candidate event of the uniformised chain, accepted or not.

Numbers on the JSON line
  value     whole-job events/s, inputs resident in HBM, timed per step with CUDA events on the
            launching stream (L2 flushed between steps), max over ranks.
  e2e       the same metric through the public API `sample_many_runs` with HOST numpy buffers:
            graph lookup, H2D of the inputs, kernels, D2H of the [R, T, 3] float64 result.
  roofline  algorithmic bytes (14 B per imitation event + 1 B per noise event + 1 B
            per accepted write) / step time, against the measured HBM copy bandwidth.
  cpu_baseline  the oracle port of the reference loop (oracle/, plain C, pthreads over runs) timed
            on this box's host cores on a bounded sample of the same workload.

`--impl reference` times that CPU implementation alone (the reference is Python + numba with no
compiled sources to build here; oracle/ is its bit-exact C restatement).
Multi-GPU: one process per GPU (torchrun), runs sharded by index, no data-path collective ("weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent-update events/sec (CNVM, ER N=1e5, ensemble)"
N_AGENTS = 100_000
RUNS_PER_GPU = 10_000
T_MAX = 100.0  # H runs to t_max = 100 with 101 output times (2.6e7 events per realization)
NUM_T = 101
WORKLOAD = (
    "H: CNVM M=3 on ER N=1e5 <d>=10 (fast_gnp seed 12345, nnz 998782), fixture rates r_imit=2 r_noise=0.6, "
    "alpha=1, t_max=100, 101 output times, OpinionShares(3) in-kernel, 1e4 realizations per GPU"
)


def measured_peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as fh:
                return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_capture() -> dict:
    """Numbers of the committed ncu capture of the dominant kernel (profiles/traffic.json), or {}."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as fh:
                return json.load(fh)
        except Exception:
            pass
    return {}


def ncu_traffic_per_launch():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    return ncu_capture().get("cnvm_native_kernel_dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = (
        "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", os.environ.get("SPONET_BENCH_SMI_MS", "200")], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin: float = 0.0, t_end: float = float("inf")) -> dict:
        """Summary of the samples taken inside [t_begin, t_end] (perf_counter clock)."""
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if not (t_begin <= ts <= t_end):
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(num_runs: int, threads: int, seed: int):
    """Runs `num_runs` realizations of H on `threads` host threads.  Returns (events, seconds)."""
    import oracle
    from sponet_b200 import workloads

    ctx = cpu_reference_run.__dict__
    if "setup" not in ctx:
        params, x0, _ = workloads.headline_cnvm(N_AGENTS)
        row_ptr, col = oracle.neighbor_list_to_csr(params.network)
        ctx["setup"] = (params, x0, row_ptr, col)
        oracle.lib()
    params, x0, row_ptr, col = ctx["setup"]
    t_eval = np.linspace(0, T_MAX, NUM_T)
    da = np.ones(N_AGENTS)  # alpha = 1
    t0 = time.perf_counter()
    _, events = oracle.sample_many_runs(
        "cnvm", x0, num_runs, t_eval, 3, row_ptr=row_ptr, col=col, r_imit=params.r_imit, r_noise=params.r_noise,
        prob_imit=params.prob_imit, prob_noise=params.prob_noise, degree_alpha=da, seed=seed, num_threads=threads)
    return int(events), time.perf_counter() - t0


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


REF_DIR = os.path.join(ROOT, "baseline", "_ref")  # tools/install_reference.sh: the unmodified reference (git-ignored)


def numba_reference_available():
    """(True, "") if the reference package installed under baseline/_ref can be imported with its numba path."""
    if not os.path.isdir(os.path.join(REF_DIR, "sponet")):
        return False, "baseline/_ref/sponet missing (tools/install_reference.sh was not run in the build container)"
    try:
        import numba  # noqa: F401
    except Exception as e:  # pragma: no cover - image without numba
        return False, f"numba not importable: {e}"
    return True, ""


class NumbaReference:
    """The UNMODIFIED reference (lueckem/SPoNet v3.0.0 installed under baseline/_ref) on workload H through its own
    public API: `sponet.sample_many_runs(params, x_init, t_max, t_eval, R, n_jobs=os.cpu_count(),
    collective_variable=OpinionShares(3))` (sponet/multiprocessing.py:18-113: a ProcessPoolExecutor of numba loops).
    Events are counted as R * t_max * lambda (the expected iterations of the reference while-loop; the reference
    does not report them; the Poisson spread is < 1e-4 relative at these sizes)."""

    def __init__(self):
        os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join("/tmp", f"numba_cache_{os.getuid()}"))
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import sponet  # the reference
        from sponet_b200 import workloads

        self.sponet = sponet
        g = workloads.er_network(N_AGENTS, 10.0, 12345)
        self.params = sponet.CNVMParameters(num_opinions=3, network=g, r=workloads.R3, r_tilde=workloads.RT3, alpha=1.0)
        self.x0 = np.random.default_rng(0).integers(0, 3, N_AGENTS).astype(np.uint8)
        self.cv = sponet.OpinionShares(3)
        self.rate = self.params.r_imit * N_AGENTS + self.params.r_noise * N_AGENTS  # alpha = 1: sum_i d_i^0 = N
        self.jobs = host_threads()

    def call(self, num_runs: int, t_max: float, seed: int) -> float:
        """Wall seconds of one sample_many_runs call (process pool spawn, pickling and JIT cache loads included)."""
        t0 = time.perf_counter()
        self.sponet.sample_many_runs(self.params, self.x0, t_max, NUM_T, num_runs, n_jobs=self.jobs,
                                     collective_variable=self.cv, rng=np.random.default_rng(seed))
        return time.perf_counter() - t0

    def spawn_cost(self) -> float:
        """Fixed cost of a call: the same call with (almost) no events (~5 s at N = 1e5)."""
        return min(self.call(self.jobs, 1e-4, 90 + i) for i in range(2))

    def measure(self, runs_per_thread: int, seed: int):
        """One bounded sample of H: (events, wall seconds of the call)."""
        runs = runs_per_thread * self.jobs
        sec = self.call(runs, T_MAX, seed)
        return runs * T_MAX * self.rate, sec, runs


def run_reference_arm(args) -> None:
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores.

    Preferred: the unmodified numba + multiprocessing reference from baseline/_ref (kind "reference").  `value` is its
    steady throughput, i.e. with the fixed process-pool / pickling / JIT-load cost of a call (measured by an empty
    call) subtracted -- the conservative choice for the ratio; the throughput of the whole call is reported beside
    it.  If the reference cannot run on this box, the C port of its loops (oracle/) is timed instead (kind "port")
    and the line says why."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    ok, why = numba_reference_available()
    port = None
    if ok:
        try:
            ref = NumbaReference()
            ref.call(threads, 0.05, 1)  # JIT compile + cache
            spawn = ref.spawn_cost()
            for w in range(min(args.warmup, 1)):
                ref.measure(1, 100 + w)
            tot_ev, tot_sec, runs = 0.0, 0.0, 0
            for k in range(args.steps):
                ev, sec, runs = ref.measure(1, 1000 + k)
                tot_ev += ev
                tot_sec += sec
            steady = tot_ev / max(tot_sec - args.steps * spawn, 1e-9)
            whole = tot_ev / tot_sec
            value, kind = steady, "reference"
            sample = (f"{runs} realizations of H (t_max=100, T=101) per step through sponet.sample_many_runs(n_jobs={threads}) "
                      f"of the unmodified reference; fixed cost of a call {spawn:.2f} s (process pool, pickling, JIT cache) subtracted")
            extra = {"value_whole_call": whole, "spawn_s_per_call": spawn, "numba": True}
            arm = "unmodified reference (baseline/_ref): numba loops + ProcessPoolExecutor, sponet/multiprocessing.py:85-109"
        except Exception as e:  # the reference is installed but does not run here
            ok, why = False, f"reference failed: {type(e).__name__}: {e}"
    if not ok:
        # calibrate: one run per thread, then size each step for ~8 s
        ev, sec = cpu_reference_run(threads, threads, 1)
        runs_per_step = max(threads, int(threads * max(1.0, 8.0 / max(sec, 1e-3))))
        for w in range(args.warmup):
            cpu_reference_run(threads, threads, 100 + w)
        tot_ev, tot_sec = 0, 0.0
        for k in range(args.steps):
            ev, sec = cpu_reference_run(runs_per_step, threads, 1000 + k)
            tot_ev += ev
            tot_sec += sec
        value, kind = tot_ev / tot_sec, "port"
        sample = f"{runs_per_step} realizations of H per step on {threads} threads"
        extra = {"numba": False, "numba_unavailable": why}
        arm = "oracle port of sponet/cnvm/model.py:239-308 (C, pthreads over runs)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_sec / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_arm": arm},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": threads, "kind": kind, "sample": sample, **extra},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    from sponet_b200 import CNVM, OpinionShares, engine, sample_many_runs, workloads
    from sponet_b200 import multiprocessing as smp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    params, x0, rate = workloads.headline_cnvm(N_AGENTS)
    model = CNVM(params)
    graph = model._graph()
    struct, keep = model._struct()
    t_eval = np.linspace(0, T_MAX, NUM_T)
    spec = (np.arange(3, dtype=np.int32), 1.0)
    d_x0 = torch.as_tensor(x0[None, :], device="cuda")
    R, R_total = RUNS_PER_GPU, RUNS_PER_GPU * world
    rb, re = rank * R, (rank + 1) * R
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step(seed, counters=None):
        return engine.run_native("cnvm", graph, N_AGENTS, struct, d_x0, t_eval, seed, R_total, rb, re,
                                 cv_spec=spec, want_cv=True, counters=counters)

    # the sampler starts before the warm-up so that the timed steps follow the warm-up without an idle gap
    # (an idle B200 drops its clocks within ~100 ms and the next launch pays the ramp); only the samples
    # taken inside the timed region are reported
    sampler = ClockSampler(local_rank)
    sampler.start()
    for w in range(args.warmup):
        step(10 + w)
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    step_ms = []
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # evict L2 between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(1000 + k, counters)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    clocks = sampler.stop(wall0, wall1)

    c = counters.cpu().numpy().astype(np.float64)
    stats = torch.tensor([sum(step_ms), c[0], c[1]], dtype=torch.float64, device="cuda")
    tmax = stats[:1].clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    total_ms = float(tmax.item())
    events, accepted = float(stats[1].item()), float(stats[2].item())
    value = events / (total_ms * 1e-3)

    # ---- e2e: public API, host buffers in, host array out ----
    cv = OpinionShares(3)
    rng = np.random.default_rng(7)
    if world > 1:
        # each rank drives its own 1e4 runs through the public API (no gather: weak scaling)
        smp_dist = smp._dist
        smp._dist = lambda: (0, 1)
    e2e_steps = max(1, min(args.steps, 5))
    sample_many_runs(params, x0, T_MAX, NUM_T, R, collective_variable=cv, rng=rng)  # warm (graph upload, caches)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, x = sample_many_runs(params, x0, T_MAX, NUM_T, R, collective_variable=cv, rng=rng)
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        smp._dist = smp_dist
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_events = events / max(args.steps, 1) * e2e_steps  # same expected work per step (Poisson counts)
    e2e_value = e2e_events / float(e2e_t.item())
    h2d = int(x0.nbytes + t_eval.nbytes + 2 * 9 * 4 + 3 * 4)
    d2h = int(x.nbytes)

    if rank == 0:
        p_acc = accepted / max(events, 1.0)
        p_noise = params.r_noise * N_AGENTS / rate
        bytes_per_event = (1.0 - p_noise) * 14.0 + p_noise * 1.0 + p_acc * 1.0
        sector_bytes = (1.0 - p_noise) * 4 * 32.0 + p_noise * 32.0
        peak, peak_src = measured_peaks()
        per_gpu_events_per_s = value / world
        achieved = per_gpu_events_per_s * bytes_per_event / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 (2-bit packed states, u32 Philox draws)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"runs sharded over {world} GPU(s), no collective",
                       "l2": "flushed between timed steps (256 MiB write)", "events_per_step_per_gpu": events / max(args.steps, 1) / world,
                       "accepted_fraction": p_acc},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "sponet_b200.sample_many_runs(params, x_init, 100.0, 101, 10000, collective_variable=OpinionShares(3))",
                    "steps": e2e_steps},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(), "traffic_source": ncu_capture().get("source"),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_event": bytes_per_event,
                         # SURVEY.md 8(d) context figure: what an HBM-resident design would move, one 32 B sector per
                         # independent random access (4 per imitation event, 1 per noise event)
                         "sector_granular": {"bytes_per_event": sector_bytes, "achieved": per_gpu_events_per_s * sector_bytes / 1e9,
                                             "frac": per_gpu_events_per_s * sector_bytes / 1e9 / peak},
                         "kernel": "cnvm_native_kernel<2,smem,nbtab,noalias,ring>",
                         "ncu": {k: v for k, v in ncu_capture().items() if k.endswith("_pct") or k.startswith("warp_")},
                         "note": "chain states live in shared memory and the CSR in L2, so DRAM traffic is far below the algorithmic bytes by design; the kernel is issue/latency bound (DESIGN.md section 5)"},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            cpu_reference_run(threads, threads, 1)  # warm + calibrate
            ev, sec = cpu_reference_run(threads, threads, 2)
            runs = max(threads, int(threads * max(1.0, 12.0 / max(sec, 1e-3))))
            ev, sec = cpu_reference_run(runs, threads, 3)
            line["cpu_baseline"] = {"value": ev / sec, "unit": "events/s", "cores": threads, "kind": "port",
                                    "sample": f"{runs} realizations of H (t_max=100, T=101) on {threads} host threads, {sec:.1f} s"}
        print(json.dumps(line), flush=True)
    del keep
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# --all-configs: every BASELINE.json configuration shape on one GPU next to the CPU arm
# (writes gpurun_out/configs.json; summarised in profiles/r1_configs.md).  Not part of the default run.
# ---------------------------------------------------------------------------------------------
def _gpu_rate(kind, model, struct, graph, n, x0, t_eval, runs, spec, counts_init=None, reps=3):
    import torch

    from sponet_b200 import engine

    d_x0 = None if x0 is None else torch.as_tensor(x0[None, :], device="cuda")
    d_c0 = None if counts_init is None else torch.as_tensor(counts_init, device="cuda")
    best = 0.0
    # warm-up: an idle B200 drops its clocks (the CPU legs run in between); keep the GPU busy for ~0.7 s first
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.7:
        engine.run_native(kind, graph, n, struct, d_x0, t_eval, 50, runs, 0, runs, cv_spec=spec, want_cv=True,
                          counts_init=d_c0)
        torch.cuda.synchronize()
    for rep in range(reps):
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        engine.run_native(kind, graph, n, struct, d_x0, t_eval, 100 + rep, runs, 0, runs, cv_spec=spec, want_cv=True,
                          counters=counters, counts_init=d_c0)
        e1.record()
        torch.cuda.synchronize()
        ev = float(counters[0].item())
        best = max(best, ev / (e0.elapsed_time(e1) * 1e-3))
    return best, ev


def _cpu_rate(model_name, x0, t_eval, M, threads, target_s=8.0, **kw):
    import oracle

    _, ev = oracle.sample_many_runs(model_name, x0, threads, t_eval, M, num_threads=threads, seed=1, **kw)  # warm
    t0 = time.perf_counter()
    _, ev = oracle.sample_many_runs(model_name, x0, threads, t_eval, M, num_threads=threads, seed=2, **kw)
    per = time.perf_counter() - t0
    runs = max(threads, int(threads * max(1.0, target_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    _, ev = oracle.sample_many_runs(model_name, x0, runs, t_eval, M, num_threads=threads, seed=3, **kw)
    sec = time.perf_counter() - t0
    return ev / sec, runs, sec


def run_all_configs():
    import torch

    import oracle  # CPU baseline leg
    from sponet_b200 import CNTM, CNVM, OpinionShares, engine, sample_moments, workloads

    threads = host_threads()
    out = []

    # config 1: CNVM M=2 complete N=1000, t_max=100, T=101 (count-only kernel; 1e5 runs to fill the GPU)
    params, x0, rate = workloads.complete_cnvm(1000)
    model = CNVM(params)
    struct, keep = model._struct()
    t_eval = np.linspace(0, 100.0, 101)
    spec = (np.arange(2, dtype=np.int32), 1.0)
    c0 = np.bincount(x0, minlength=2).astype(np.int64)[None, :]
    g, _ = _gpu_rate("cnvm_counts", model, struct, None, 1000, None, t_eval, 100_000, spec, counts_init=c0)
    g1k, _ = _gpu_rate("cnvm_counts", model, struct, None, 1000, None, t_eval, 1000, spec, counts_init=c0)
    ga, _ = _gpu_rate("cnvm", model, struct, None, 1000, x0, t_eval, 100_000, spec)
    c, runs, sec = _cpu_rate("cnvm_complete", x0, t_eval, 2, threads, r_imit=params.r_imit, r_noise=params.r_noise,
                            prob_imit=params.prob_imit, prob_noise=params.prob_noise, degree_alpha=model.degree_alpha)
    out.append(dict(config="1: CNVM M=2 complete N=1000, t_max=100, T=101", gpu_events_per_s=g,
                    note=f"count-only kernel, 1e5 runs; 1e3 runs (as in BASELINE): {g1k:.3g}; agent-level kernel, 1e5 runs: {ga:.3g}",
                    cpu_events_per_s=c, cpu_threads=threads, cpu_sample=f"{runs} runs, {sec:.1f} s"))

    # config 2: CNVM M=3 ER N=1e4, R=1e4, T=100
    params, x0, rate = workloads.headline_cnvm(10_000)
    model = CNVM(params)
    struct, keep = model._struct()
    t_eval = np.linspace(0, 10.0, 100)
    spec = (np.arange(3, dtype=np.int32), 1.0)
    g, _ = _gpu_rate("cnvm", model, struct, model._graph(), 10_000, x0, t_eval, 10_000, spec)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    c, runs, sec = _cpu_rate("cnvm", x0, t_eval, 3, threads, row_ptr=rp, col=col, r_imit=params.r_imit,
                            r_noise=params.r_noise, prob_imit=params.prob_imit, prob_noise=params.prob_noise,
                            degree_alpha=np.ones(10_000))
    out.append(dict(config="2: CNVM M=3 ER N=1e4 <d>=10, 1e4 realizations, 100 outputs (t_max=10)", gpu_events_per_s=g,
                    cpu_events_per_s=c, cpu_threads=threads, cpu_sample=f"{runs} runs, {sec:.1f} s"))

    # config 3: CNTM BA N=1e5 m=5, R=1e4
    params, x0, rate = workloads.ba_cntm(100_000)
    model = CNTM(params)
    t_eval = np.linspace(0, 2.0, 11)
    spec = (np.arange(2, dtype=np.int32), 1.0)
    g, _ = _gpu_rate("cntm", model, model._struct(), model._graph(), 100_000, x0, t_eval, 10_000, spec)
    rp, col = oracle.neighbor_list_to_csr(model.neighbor_list)
    c, runs, sec = _cpu_rate("cntm", x0, t_eval, 2, threads, row_ptr=rp, col=col, next_event_rate=model.next_event_rate,
                            noise_prob=model.noise_prob, threshold_01=0.5, threshold_10=0.5)
    out.append(dict(config="3: CNTM BA N=1e5 m=5, r=1 r~=0.2 thr 0.5/0.5, 1e4 realizations (t_max=2 sample)",
                    gpu_events_per_s=g, cpu_events_per_s=c, cpu_threads=threads, cpu_sample=f"{runs} runs, {sec:.1f} s"))

    # config 4: SIS-as-CNVM WS N=1e6 (per-GPU share of 1e5 realizations: timed on 2368 = 16 per SM)
    params, x0, rate = workloads.sis_ws(1_000_000)
    model = CNVM(params)
    struct, keep = model._struct()
    t_eval = np.linspace(0, 1.0, 11)
    spec = (np.array([1], dtype=np.int32), 1e6)
    g, _ = _gpu_rate("cnvm", model, struct, model._graph(), 1_000_000, x0, t_eval, 2368, spec)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    c, runs, sec = _cpu_rate("cnvm", x0, t_eval, 2, threads, target_s=6.0, row_ptr=rp, col=col, r_imit=params.r_imit,
                            r_noise=params.r_noise, prob_imit=params.prob_imit, prob_noise=params.prob_noise,
                            degree_alpha=np.ones(1_000_000))
    out.append(dict(config="4: SIS-as-CNVM WS N=1e6 k=10 p=0.1, lambda=1 (2368 realizations, t_max=1 sample)",
                    gpu_events_per_s=g, cpu_events_per_s=c, cpu_threads=threads, cpu_sample=f"{runs} runs, {sec:.1f} s"))

    # config 5: one (r, r~) point of the sample_moments sweep on the H graph, fused integer moments (wall clock, host API)
    from sponet_b200 import CNVMParameters

    hp, hx0, _ = workloads.headline_cnvm(100_000)
    p5 = CNVMParameters(num_opinions=2, network=hp.network, r=1.0, r_tilde=0.02)
    x5 = (hx0 % 2).astype(np.uint8)
    cv = OpinionShares(2, normalize=True)
    sample_moments(p5, x5, 1.0, 11, 2000, rel_tol=1e9, collective_variable=cv, rng=np.random.default_rng(0))  # warm
    t0 = time.perf_counter()
    _, mean, var, n = sample_moments(p5, x5, 10.0, 101, 10_000, rel_tol=1e9, collective_variable=cv, rng=np.random.default_rng(1))
    sec = time.perf_counter() - t0
    ev = 10_000 * 10.0 * (p5.r_imit + p5.r_noise) * 100_000
    out.append(dict(config="5: one point of the (r, r~) sweep: sample_moments, M=2, ER N=1e5, 1e4 runs, t_max=10, T=101",
                    gpu_events_per_s=ev / sec, note=f"wall clock through sample_moments (host API), {sec:.2f} s per point",
                    cpu_events_per_s=None))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/configs.json", "w") as fh:
        json.dump(out, fh, indent=1)
    for o in out:
        print(json.dumps(o))
    del keep



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-configs", action="store_true", help="measure every BASELINE config shape (one GPU)")
    args = ap.parse_args()
    if args.all_configs:
        run_all_configs()
        return
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        port = str(29500 + os.getpid() % 2000)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", port,
                                   os.path.abspath(__file__), *sys.argv[1:]])
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
