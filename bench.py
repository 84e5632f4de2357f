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
def numba_baseline_subprocess():
    """cpu_baseline of the GPU line: the unmodified numba reference in a SUBPROCESS (its process pool forks; this
    process holds a CUDA context), one bounded step.  Returns the dict of that run's `cpu_baseline`, or a reason."""
    ok, why = numba_reference_available()
    if not ok:
        return {"unavailable": why}
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(k, None)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, timeout=900, env=env)
        line = json.loads(r.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}


def sharded_equals_single(params, x0, graph, struct, world: int, rank: int) -> bool:
    """Self-check before the timed region: a small ensemble through the REAL (distributed) public API -- seed
    broadcast, runs split over the ranks, all-gather -- equals the same ensemble computed by ONE launch on this
    GPU, bit for bit.  At N = 1 the sharding is emulated by two launches over complementary run ranges."""
    from sponet_b200 import OpinionShares, engine, sample_many_runs

    R, T, t_max = 96, 6, 0.5
    t_eval = np.linspace(0, t_max, T)
    spec = (np.arange(3, dtype=np.int32), 1.0)
    _, c_api = sample_many_runs(params, x0, t_max, T, R, collective_variable=OpinionShares(3), rng=np.random.default_rng(5))
    seed = engine.draw_seed(np.random.default_rng(5))  # the key sample_many_runs drew (rank 0's under torch.distributed)
    _, one = engine.run_native("cnvm", graph, N_AGENTS, struct, x0[None, :], t_eval, seed, R, 0, R, cv_spec=spec, want_cv=True)
    ok = c_api.shape == (R, T, 3) and np.array_equal(c_api, one[0])
    if world == 1:
        _, lo = engine.run_native("cnvm", graph, N_AGENTS, struct, x0[None, :], t_eval, seed, R, 0, 40, cv_spec=spec, want_cv=True)
        _, hi = engine.run_native("cnvm", graph, N_AGENTS, struct, x0[None, :], t_eval, seed, R, 40, R, cv_spec=spec, want_cv=True)
        ok = ok and np.array_equal(np.concatenate([lo[0], hi[0]]), one[0])
    return bool(ok)


def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    from sponet_b200 import CNVM, OpinionShares, engine, sample_many_runs, workloads
    from sponet_b200.multiprocessing import _split_runs

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
    strong = args.scaling == "strong"
    R_total = RUNS_PER_GPU if strong else RUNS_PER_GPU * world  # strong: SURVEY.md 8(d) H has 1e4 realizations in total
    bounds = np.concatenate(([0], np.cumsum(_split_runs(R_total, world))))
    rb, re = int(bounds[rank]), int(bounds[rank + 1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step(seed, counters=None):
        return engine.run_native("cnvm", graph, N_AGENTS, struct, d_x0, t_eval, seed, R_total, rb, re,
                                 cv_spec=spec, want_cv=True, counters=counters)

    check = sharded_equals_single(params, x0, graph, struct, world, rank)
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
    stats = torch.tensor([sum(step_ms), c[0], c[1], float(check)], dtype=torch.float64, device="cuda")
    tmax = stats[:1].clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    total_ms = float(tmax.item())
    events, accepted = float(stats[1].item()), float(stats[2].item())
    check_all = float(stats[3].item()) == float(world)
    value = events / (total_ms * 1e-3)

    # ---- e2e: the public API with host buffers in and a host array out, exactly as a user calls it.  Under
    # torch.distributed that is the real sharded path: seed broadcast, this rank's share of the runs, NCCL
    # all-gather of the per-rank slabs, device-to-host copy of the full [R_total, T, 3] array on every rank.
    cv = OpinionShares(3)
    rng = np.random.default_rng(7)
    e2e_steps = max(1, min(args.steps, 5))
    sample_many_runs(params, x0, T_MAX, NUM_T, R_total, collective_variable=cv, rng=rng)  # warm (graph upload, caches)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, x = sample_many_runs(params, x0, T_MAX, NUM_T, R_total, collective_variable=cv, rng=rng)
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    assert x.shape == (R_total, NUM_T, 3)
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
        workload = WORKLOAD if not strong else WORKLOAD.replace("1e4 realizations per GPU", "1e4 realizations in total (strong scaling)")
        line = {
            "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u8 (2-bit packed states, u32 Philox draws)",
            "data": "synthetic",
            "config": {"workload": workload,
                       "parallelism": f"runs sharded over {world} GPU(s); value: no collective in the timed region; e2e: "
                                      + ("seed broadcast + NCCL all-gather of the per-rank CV slabs" if world > 1 else "single GPU"),
                       "l2": "flushed between timed steps (256 MiB write)", "events_per_step_per_gpu": events / max(args.steps, 1) / world,
                       "accepted_fraction": p_acc},
            "clocks": clocks,
            "sharded_equals_single": check_all,
            "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": f"sponet_b200.sample_many_runs(params, x_init, 100.0, 101, {R_total}, collective_variable=OpinionShares(3))"
                           + (" under torch.distributed (unpatched: broadcast, shard, all-gather)" if world > 1 else ""),
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
                         "kernel": "cnvm_native_kernel<2,smem,nbtab,noalias,coop>",
                         "ncu": {k: v for k, v in ncu_capture().items() if k.endswith("_pct") or k.startswith("warp_")},
                         "note": "chain states live in shared memory and the CSR in L2, so DRAM traffic is far below the algorithmic bytes by design; the kernel is issue/latency bound (DESIGN.md section 5)"},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            nb = numba_baseline_subprocess()
            cpu_reference_run(threads, threads, 1)  # warm + calibrate
            ev, sec = cpu_reference_run(threads, threads, 2)
            runs = max(threads, int(threads * max(1.0, 8.0 / max(sec, 1e-3))))
            ev, sec = cpu_reference_run(runs, threads, 3)
            port = {"value": ev / sec, "unit": "events/s", "cores": threads, "kind": "port",
                    "sample": f"{runs} realizations of H (t_max=100, T=101) on {threads} host threads, {sec:.1f} s"}
            if "value" in nb:
                line["cpu_baseline"] = nb
                line["cpu_baseline_port"] = port
            else:  # the numba reference could not run on this box: the C port of its loops stands in
                port["numba_unavailable"] = nb.get("unavailable")
                line["cpu_baseline"] = port
        if world == 1 and not args.no_configs:
            try:
                line["configs"] = other_configs(peak)
            except Exception as e:  # the headline line must not depend on the side measurements
                line["configs"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    del keep
    if world > 1:
        dist.destroy_process_group()


def other_configs(peak_gbs: float) -> list:
    """The other single-GPU BASELINE.json configurations, measured in the same run (device-timed, best of 3 after a
    warm-up, inputs resident), each with its algorithmic bytes per event, the fraction of the HBM
    roofline and the C port of the reference loop on this box's host threads beside it.
      2: CNVM M=3, ER N=1e4, 1e4 realizations, 100 output times (t_max = 10)
      3: CNTM on Barabasi-Albert N=1e5 m=5, 1e4 realizations, t_max = 20, 101 output times
      4: SIS-as-CNVM on Watts-Strogatz N=1e6 (this GPU's share: 2368 realizations, t_max = 1)"""
    import oracle  # CPU baseline legs only
    from sponet_b200 import CNTM, CNVM, workloads

    threads = host_threads()
    out = []

    def entry(name, kernel, g_rate, events, accepted, bytes_fn, cpu):
        p_acc = accepted / max(events, 1.0)
        bpe = bytes_fn(p_acc)
        out.append({"config": name, "kernel": kernel, "events_per_s": g_rate, "algorithmic_bytes_per_event": bpe,
                    "roofline_frac": g_rate * bpe / 1e9 / peak_gbs, "cpu_port_events_per_s": cpu[0], "cpu_threads": threads,
                    "cpu_sample": f"{cpu[1]} runs, {cpu[2]:.1f} s"})

    # config 2
    params, x0, rate = workloads.headline_cnvm(10_000)
    model = CNVM(params)
    struct, keep = model._struct()
    t_eval = np.linspace(0, 10.0, 100)
    spec = (np.arange(3, dtype=np.int32), 1.0)
    g, ev, acc = _gpu_rate("cnvm", model, struct, model._graph(), 10_000, x0, t_eval, 10_000, spec)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    cpu = _cpu_rate("cnvm", x0, t_eval, 3, threads, target_s=3.0, row_ptr=rp, col=col, r_imit=params.r_imit,
                    r_noise=params.r_noise, prob_imit=params.prob_imit, prob_noise=params.prob_noise, degree_alpha=np.ones(10_000))
    pn = params.r_noise * 10_000 / rate
    entry("2: CNVM M=3 ER N=1e4 <d>=10, 1e4 realizations, 100 outputs (t_max=10)", "cnvm_native_kernel<2,smem,nbtab,noalias>",
          g, ev, acc, lambda pa: (1 - pn) * 14.0 + pn + pa, cpu)

    # config 3
    params, x0, rate = workloads.ba_cntm(100_000)
    model = CNTM(params)
    t_eval = np.linspace(0, 20.0, 101)
    spec2 = (np.arange(2, dtype=np.int32), 1.0)
    g, ev, acc = _gpu_rate("cntm", model, model._struct(), model._graph(), 100_000, x0, t_eval, 10_000, spec2, reps=2, warm_s=0.0)
    rp, col = oracle.neighbor_list_to_csr(model.neighbor_list)
    cpu = _cpu_rate("cntm", x0, np.linspace(0, 2.0, 11), 2, threads, target_s=3.0, row_ptr=rp, col=col,
                    next_event_rate=model.next_event_rate, noise_prob=model.noise_prob, threshold_01=0.5, threshold_10=0.5)
    mean_deg = float(len(col)) / 100_000
    entry("3: CNTM BA N=1e5 m=5, r=1 r~=0.2 thr 0.5/0.5, 1e4 realizations, t_max=20, 101 outputs", "cntm_native_kernel<smem>",
          g, ev, acc, lambda pa: 1.0 + (1 - model.noise_prob) * (8.0 + 5.0 * mean_deg) + pa, cpu)

    # config 4 (this GPU's share)
    params, x0, rate = workloads.sis_ws(1_000_000)
    model = CNVM(params)
    struct, keep = model._struct()
    t_eval = np.linspace(0, 1.0, 11)
    spec1 = (np.array([1], dtype=np.int32), 1e6)
    g, ev, acc = _gpu_rate("cnvm", model, struct, model._graph(), 1_000_000, x0, t_eval, 2368, spec1, reps=2)
    rp, col = oracle.neighbor_list_to_csr(params.network)
    cpu = _cpu_rate("cnvm", x0, t_eval, 2, threads, target_s=3.0, row_ptr=rp, col=col, r_imit=params.r_imit,
                    r_noise=params.r_noise, prob_imit=params.prob_imit, prob_noise=params.prob_noise, degree_alpha=np.ones(1_000_000))
    pn = params.r_noise / (params.r_noise + params.r_imit)
    entry("4: SIS-as-CNVM WS N=1e6 k=10 p=0.1, lambda=1 (2368 realizations, t_max=1: one GPU's sample of the 1e5)",
          "cnvm_native_kernel<1,smem,nbtab,noalias,coop16>", g, ev, acc, lambda pa: (1 - pn) * 14.0 + pn + pa, cpu)
    del keep
    return out


# ---------------------------------------------------------------------------------------------
# --workload sweep: BASELINE config 5, the parameter sweep with the moments all-reduced over the GPUs
# ---------------------------------------------------------------------------------------------
def run_sweep(args) -> None:
    """256 (r, r~) points on the H graph (M = 2), `--sweep-runs` realizations per point and GPU-eighth (1e5 on the
    8-GPU box: weak scaling in the number of runs per point), t_max = 1, 11 output times, OpinionShares moments.
    One launch per GPU covers its share of the points; sum c / sum c^2 stay on the device and meet in ONE NCCL
    all-reduce, which is INSIDE the timed region."""
    import torch
    import torch.distributed as dist

    from sponet_b200 import CNVMParameters, OpinionShares, engine, sample_moments_sweep, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hp, hx0, _ = workloads.headline_cnvm(N_AGENTS)
    grid_r = np.linspace(0.5, 2.0, 16)
    grid_rt = np.linspace(0.005, 0.05, 16)
    plist = [CNVMParameters(num_opinions=2, network=hp.network, r=float(r), r_tilde=float(rt)) for r in grid_r for rt in grid_rt]
    P = len(plist)
    x0 = (hx0 % 2).astype(np.uint8)
    runs = max(1, int(args.sweep_runs * world / 8))
    t_max, T = 1.0, 11
    cv = OpinionShares(2, normalize=True)
    ev_per_step = sum(runs * t_max * (q.r_imit + q.r_noise) * N_AGENTS for q in plist)
    rng = np.random.default_rng(11)
    sample_moments_sweep(plist[:8], x0, 0.1, 3, 8 * world, cv, rng=rng)  # warm (graph upload, caches)

    # bit-equality of one sweep point with a stand-alone run of its parameters (small, before the timed region)
    t_s, m_s, v_s = sample_moments_sweep(plist[:16], x0, 0.2, 3, 64, cv, rng=np.random.default_rng(3))
    from sponet_b200 import CNVM
    pm = CNVM(plist[5])
    struct, keep = pm._struct()
    seed = engine.draw_seed(np.random.default_rng(3))
    sums = (np.zeros((1, 3, 2), dtype=np.int64), np.zeros((1, 3, 2), dtype=np.int64))
    engine.run_native("cnvm", pm._graph(), N_AGENTS, struct, x0[None, :], t_s, seed, 64, 0, 64, cv_spec=cv._kernel_spec(N_AGENTS),
                      sums=sums, chain_offset=5 * 64)
    point_ok = bool(np.array_equal(sums[0][0] / 64 / N_AGENTS, m_s[5]))
    del keep

    stats_acc = {"kernel_s": 0.0, "allreduce_s": 0.0}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        st = {}
        t, mean, var = sample_moments_sweep(plist, x0, t_max, T, runs, cv, rng=rng, stats=st)
        for key in stats_acc:
            stats_acc[key] += st.get(key, 0.0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sec = time.perf_counter() - t0
    tt = torch.tensor([sec, stats_acc["kernel_s"], stats_acc["allreduce_s"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec, ks, ars = (float(v) for v in tt.cpu())
        value = ev_per_step * args.steps / sec
        line = {"metric": "agent-update events/sec (CNVM parameter sweep, ER N=1e5, moments all-reduced)", "value": value,
                "unit": "events/s", "n_gpus": world, "steps": args.steps, "warmup": 1, "ms_per_step": 1e3 * sec / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (1-bit packed states), int64 moment sums",
                "data": "synthetic",
                "config": {"workload": f"config 5: 16 x 16 grid of (r, r~) in [0.5,2] x [0.005,0.05], M=2 on the H graph (ER N=1e5), "
                                       f"{runs} realizations per point ({args.sweep_runs} on 8 GPUs), t_max=1, 11 output times, "
                                       "OpinionShares moments", "points": P,
                           "parallelism": f"points dealt to {world} GPU(s), one launch per GPU, one all-reduce of [2, {P}, {T}, 2] int64"},
                "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": int(x0.nbytes), "d2h_bytes_per_step": int(2 * P * T * 2 * 8),
                        "api": "sponet_b200.sample_moments_sweep(params_list, x_init, 1.0, 11, runs, OpinionShares(2, normalize=True))"},
                "allreduce": {"seconds_per_step": ars / args.steps, "share_of_step": ars / sec if sec > 0 else 0.0,
                              "bytes": int(2 * P * T * 2 * 8), "inside_timed_region": True},
                "kernel_seconds_per_step": ks / args.steps,
                "sweep_point_equals_standalone": point_ok, "gpu_launches": 2 * args.steps,
                "sample_mean_first_last_point": [float(mean[0, -1, 1]), float(mean[-1, -1, 1])]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# --workload networks: H with a NEW Erdos-Renyi network for every realization
# ---------------------------------------------------------------------------------------------
# Mean degree 16 with rejection of samples that contain an isolated vertex (1 % of them): the closest workload to H that the
# reference can run.  With H's mean degree 10 a sample has ~4.5 isolates: the reference then either gives up after
# max_sample_time (RuntimeError, sponet/network_generator.py:75-78) or, with force_no_isolates=True, never returns --
# `while (j := i) == i` in _unisolate_vertices (:451) is always true.  (tools/probe_networks.py measures this repo on H's own
# degree with the repair rule: profiles/r2_kernels.md section 6.)
NET_DEGREE = 16.0
NET_P = NET_DEGREE / (N_AGENTS - 1)


def run_networks_reference(args) -> None:
    """The unmodified reference on the same workload: `sponet.sample_many_runs` with `sponet.ErdosRenyiGenerator` in the
    parameters -- every worker draws a networkx graph before every run (sponet/cnvm/model.py:95-96)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    ok, why = numba_reference_available()
    if not ok:
        print(json.dumps({"impl": "reference", "unavailable": why}), flush=True)
        return
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join("/tmp", f"numba_cache_{os.getuid()}"))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import sponet
    from sponet_b200 import workloads

    threads = host_threads()
    gen = sponet.ErdosRenyiGenerator(N_AGENTS, NET_P, max_sample_time=120, rng=np.random.default_rng(1))
    params = sponet.CNVMParameters(num_opinions=3, network_generator=gen, r=workloads.R3, r_tilde=workloads.RT3, alpha=1.0)
    x0 = np.random.default_rng(0).integers(0, 3, N_AGENTS).astype(np.uint8)
    cv = sponet.OpinionShares(3)
    rate = params.r_imit * N_AGENTS + params.r_noise * N_AGENTS

    def call(runs, t_max, seed):
        t0 = time.perf_counter()
        sponet.sample_many_runs(params, x0, t_max, NUM_T, runs, n_jobs=threads, collective_variable=cv, rng=np.random.default_rng(seed))
        return time.perf_counter() - t0

    call(threads, 0.05, 1)  # JIT compile + cache
    tot_ev, tot_sec = 0.0, 0.0
    for k in range(max(args.steps, 1)):
        tot_sec += call(threads, T_MAX, 1000 + k)
        tot_ev += threads * T_MAX * rate
    value = tot_ev / tot_sec
    line = {"impl": "reference", "metric": "agent-update events/sec (CNVM, a new ER N=1e5 network per realization)", "value": value,
            "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": 0, "ms_per_step": 1e3 * tot_sec / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": f"H's model with network_generator=ErdosRenyiGenerator({N_AGENTS}, {NET_DEGREE:g}/(N-1)): "
                                   f"{threads} realizations per step, t_max=100, T=101",
                       "reference_arm": "unmodified reference (baseline/_ref): sponet.sample_many_runs, whole call (process pool included)"},
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": threads, "kind": "reference",
                             "sample": f"{threads} realizations per step, each on its own networkx graph"},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_networks(args) -> None:
    """1e4 realizations per GPU of H's model, each on its own G(N, p) generated on the device, through the public API
    (`sample_many_runs` with `sponet_b200.ErdosRenyiGenerator`): host buffers in and out, graph generation, pooled batch
    memory and the all-gather of the result inside the timed region (wall clock, max over ranks)."""
    import torch
    import torch.distributed as dist

    from sponet_b200 import CNVMParameters, ErdosRenyiGenerator, OpinionShares, sample_many_runs, workloads
    from sponet_b200 import multiprocessing as mp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    x0 = np.random.default_rng(0).integers(0, 3, N_AGENTS).astype(np.uint8)
    cv = OpinionShares(3)
    R = RUNS_PER_GPU * world

    def params(seed):
        gen = ErdosRenyiGenerator(N_AGENTS, NET_P, rng=np.random.default_rng(seed))
        return CNVMParameters(num_opinions=3, network_generator=gen, r=workloads.R3, r_tilde=workloads.RT3, alpha=1.0)

    # self-check: the result does not depend on how the realizations are cut into blocks of graphs
    _, a = sample_many_runs(params(3), x0, 0.5, 4, 24 * world, collective_variable=cv, rng=np.random.default_rng(4))
    mp._BATCH_BYTES = int(7 * 16e6)
    _, b = sample_many_runs(params(3), x0, 0.5, 4, 24 * world, collective_variable=cv, rng=np.random.default_rng(4))
    mp._BATCH_BYTES = None
    blocks_ok = bool(np.array_equal(a, b))
    sampler = ClockSampler(local_rank)
    sampler.start()
    for w in range(args.warmup):
        sample_many_runs(params(10 + w), x0, T_MAX / 20, NUM_T, R, collective_variable=cv, rng=np.random.default_rng(w))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        _, c = sample_many_runs(params(100 + k), x0, T_MAX, NUM_T, R, collective_variable=cv, rng=np.random.default_rng(200 + k))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1)
    tt = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(tt.cpu()[0])
        rate = (workloads.R3.max() + 3 * workloads.RT3.max()) * N_AGENTS
        value = R * T_MAX * rate * args.steps / sec
        base = None
        if not args.no_cpu_baseline:
            try:
                env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
                for key in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                    env.pop(key, None)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", "networks",
                                    "--steps", "1"], capture_output=True, text=True, timeout=900, env=env)
                base = json.loads(r.stdout.strip().splitlines()[-1]).get("cpu_baseline")
            except Exception as e:
                base = {"unavailable": f"{type(e).__name__}: {e}"}
        line = {"metric": "agent-update events/sec (CNVM, a new ER N=1e5 network per realization)", "value": value, "unit": "events/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 (2-bit packed states), uint32 graph records", "data": "synthetic",
                "config": {"workload": f"H's model with network_generator=ErdosRenyiGenerator({N_AGENTS}, {NET_DEGREE:g}/(N-1)): "
                                       f"{RUNS_PER_GPU} realizations per GPU, each on its own graph generated on the device, "
                                       "t_max=100, T=101, OpinionShares",
                           "parallelism": f"realizations split over {world} GPU(s), blocks of graphs per launch"},
                "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": int(x0.nbytes), "d2h_bytes_per_step": int(c.nbytes),
                        "api": "sponet_b200.sample_many_runs(CNVMParameters(network_generator=ErdosRenyiGenerator(...)), x0, 100, 101, R, OpinionShares(3))"},
                "block_size_independent": blocks_ok, "clocks": clocks, "cpu_baseline": base,
                "sample_mean_final_shares": [float(v) for v in c[:, -1].mean(0) / N_AGENTS]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# --all-configs: every BASELINE.json configuration shape on one GPU next to the CPU arm
# (writes gpurun_out/configs.json; summarised in profiles/r1_configs.md).  Not part of the default run.
# ---------------------------------------------------------------------------------------------
def _gpu_rate(kind, model, struct, graph, n, x0, t_eval, runs, spec, counts_init=None, reps=3, warm_s=0.7):
    import torch

    from sponet_b200 import engine

    d_x0 = None if x0 is None else torch.as_tensor(x0[None, :], device="cuda")
    d_c0 = None if counts_init is None else torch.as_tensor(counts_init, device="cuda")
    best = 0.0
    # warm-up: an idle B200 drops its clocks (the CPU legs run in between); keep the GPU busy for ~0.7 s first
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < warm_s:
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
        ev, acc = float(counters[0].item()), float(counters[1].item())
        best = max(best, ev / (e0.elapsed_time(e1) * 1e-3))
    return best, ev, acc


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
    g, _, _ = _gpu_rate("cnvm_counts", model, struct, None, 1000, None, t_eval, 100_000, spec, counts_init=c0)
    g1k, _, _ = _gpu_rate("cnvm_counts", model, struct, None, 1000, None, t_eval, 1000, spec, counts_init=c0)
    ga, _, _ = _gpu_rate("cnvm", model, struct, None, 1000, x0, t_eval, 100_000, spec)
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
    g, _, _ = _gpu_rate("cnvm", model, struct, model._graph(), 10_000, x0, t_eval, 10_000, spec)
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
    g, _, _ = _gpu_rate("cntm", model, model._struct(), model._graph(), 100_000, x0, t_eval, 10_000, spec)
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
    g, _, _ = _gpu_rate("cnvm", model, struct, model._graph(), 1_000_000, x0, t_eval, 2368, spec)
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
    ap.add_argument("--no-configs", action="store_true", help="skip the side measurements of configs 2-4")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 1e4 realizations per GPU; strong: 1e4 realizations in total (SURVEY.md 8(d))")
    ap.add_argument("--workload", default="H", choices=["H", "sweep", "networks"],
                    help="sweep: BASELINE config 5 (moments all-reduced); networks: H with a new network per realization")
    ap.add_argument("--sweep-runs", type=int, default=100_000, help="realizations per sweep point on 8 GPUs (scaled by N/8)")
    ap.add_argument("--all-configs", action="store_true", help="measure every BASELINE config shape (one GPU)")
    args = ap.parse_args()
    if args.all_configs:
        run_all_configs()
        return
    if args.impl == "reference":
        if args.workload == "networks":
            run_networks_reference(args)
        else:
            run_reference_arm(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        port = str(29500 + os.getpid() % 2000)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", port,
                                   os.path.abspath(__file__), *sys.argv[1:]])
    if args.workload == "sweep":
        run_sweep(args)
        return
    if args.workload == "networks":
        run_networks(args)
        return
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
