#!/usr/bin/env python
"""Benchmark of the interior-point Newton step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--horizon 5000] [--no-cpu-baseline] [--multistart B]

A *step* is one Newton iteration of the interior-point loop
(reference: include/sleipnir/optimization/solver/interior_point.hpp:382-863) on
the cart-pole direct-transcription problem of the reference's scalability
benchmark (benchmarks/scalability/cart_pole/sleipnir.cpp, T = 5 s, dt = T/N),
N = 5000 by default, started from the benchmark's initial guess. W warm-up
iterations are followed by exactly K timed ones inside one solve (K = 200 by
default: a cart-pole solve of this size runs for hundreds of iterations); every
iteration ends with a device→host read of its scalars, so the host timestamps
taken at iteration boundaries are device-complete. The L2 is evicted before
every timed iteration (excluded from the timestamps).

Besides the contract's keys the line carries `roofline` (k_factor_tree against
the measured HBM peak, plus the derivative sweep), `cpu_baseline` (the oracle on
one host core), `e2e` (one Problem::solve() to its exit status from host
buffers, setup included) and `multistart` (B such solves in flight on each GPU
through slp::multistart; --multistart 0 skips it).

Prints ONE JSON line (see README/DESIGN for the field meanings).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Newton steps/sec (cart-pole direct transcription, interior-point)"
UNIT = "steps/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region. Uses NVML
    in-process (pynvml) — spawning nvidia-smi five times a second contends for
    the driver lock with the very cudaMalloc/memcpy calls being timed — and
    falls back to nvidia-smi when pynvml is unavailable."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown",
               0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index=0):
        self.samples = []     # (sm_mhz, max_mhz, set of reasons)
        self.gpu_index = gpu_index
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
        except Exception:
            self._nvml = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu_index])
            except Exception:
                pass
        return self.gpu_index

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._handle, n.NVML_CLOCK_SM)
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
        self.samples.append((float(sm), float(mx),
                             {name for bit, name in self.REASONS.items() if bits & bit}))

    def _sample_smi(self):
        out = subprocess.run(
            ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
             "-i", str(self._physical_index())],
            capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        c = [x.strip() for x in out.split(",")]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        self.samples.append((float(c[1]), float(c[2]),
                             {nm for nm, v in zip(names, c[5:9])
                              if v.lower().startswith("active")}))

    enabled = True

    def _loop(self):
        while self.enabled and not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.5)

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def steady_rate(trace, warmup, steps):
    """steps / (t_end[W+K-1] - t_end[W-1]) from per-iteration timestamps."""
    if len(trace) < warmup + steps:
        steps = len(trace) - warmup
    t0 = trace[warmup - 1].t_end if warmup > 0 else 0.0
    t1 = trace[warmup + steps - 1].t_end
    return steps, (t1 - t0)


def pin_to_one_core(index=0):
    """taskset for this process (the reference path runs on one
    pinned core). Returns the core, or None when the platform refuses."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        core = cores[index % len(cores)]
        os.sched_setaffinity(0, {core})
        return core
    except Exception:
        return None


def golden_iterations(horizon):
    """Iterations the reference algorithm (oracle, AMD order, default Options)
    runs to its own exit status on this workload — committed golden,
    tests/golden/make_convergence_golden.py. None when there is no golden."""
    path = os.path.join(ROOT, "tests", "golden", f"converge_cart_pole_{horizon}.npz")
    if not os.path.exists(path):
        return None, None
    import numpy as np
    g = np.load(path)
    return int(g["iterations"]), int(g["status"])


def run_cpu(horizon, steps, warmup, core_index=0):
    """The reference-style CPU path (oracle), single thread like the reference
    (it never uses more than one thread per solve), pinned to one core."""
    from oracle.pyoracle import OracleProblem, have_reference
    core = pin_to_one_core(core_index)
    backend = "reference" if have_reference() else "restated"
    t0 = time.perf_counter()
    P = OracleProblem("cart_pole", horizon, backend=backend)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    P.solve(max_iterations=warmup + steps, keep_iterates=False)
    total_s = time.perf_counter() - t0
    tr = P.trace()
    k, dt = steady_rate(tr, warmup, steps)
    P.close()
    return {"rate": k / dt, "steps": k, "loop_s": dt, "total_s": total_s,
            "build_s": build_s, "backend": backend, "core": core,
            "setup_s": build_s + max(total_s - (tr[-1].t_end if tr else 0.0), 0.0),
            "fact_per_step": sum(r.factorizations for r in tr) / len(tr),
            "iters": len(tr)}


def amd_nnz_l(sb, N, device):
    """nnz(L) of the workload's KKT system under the reference's minimum-degree
    order (SLPB_ORDER_AMD), for the roofline's second denominator."""
    P = sb.Problem("cart_pole", N)
    D = P.open_device(device)
    import numpy as np
    D.set_iterate(P.initial_guess(), np.ones(P.mi), np.zeros(P.me), np.ones(P.mi))
    D.eval_current(1)
    st = D.analyze(sb.ORDER_AMD)
    nnz = int(st.nnz_l)
    P.close_device()
    P.close()
    return nnz


def run_batch(sb, N, B, device, peak, fac_bytes, sol_bytes, fac_bytes_amd, sol_bytes_amd,
              arithmetic=0):
    """The many-instance regime (slp::multistart's data-parallel
    axis): B KKT systems of the workload — the lhs of one iterate and B − 1
    perturbed copies — factored and solved by ONE launch each (slpb_batch_factor,
    slpb_batch_solve; lane = instance, SoA [entry][32]). Same dependency chain
    as one factorisation, B× the bytes: the regime in which the LDLT is bound by
    memory traffic. Inputs (B × 1.2 MB of values) and the factor (B × 5 MB)
    exceed the L2 for B ≥ 32, so every launch streams from HBM."""
    import numpy as np
    P = sb.Problem("cart_pole", N)
    D = P.open_device(device)
    rng = np.random.default_rng(0)
    x = P.initial_guess() + 0.01 * rng.standard_normal(P.n)
    D.set_iterate(x, 0.5 + rng.random(P.mi), 0.1 * rng.standard_normal(P.me),
                  0.5 + rng.random(P.mi))
    D.eval_current(1)
    D.analyze()
    D.set_factor_arithmetic(arithmetic)
    fi = D.factor(1.0, 1e-6, True)
    D.solve(0.1, 0.99)
    kkt, rhs = D.download(sb.ARR_KKT_VAL), D.download(sb.ARR_RHS)
    sol1 = np.concatenate([D.download(sb.ARR_P_X), -D.download(sb.ARR_P_Y)])
    Bt = sb.Batch(D, B)
    for i in range(B):
        Bt.set_system(i, kkt if i == 0 else kkt * (1 + 1e-3 * rng.standard_normal(kkt.size)), rhs)
    f_ms, s_ms = [], []
    for rep in range(5):
        info = Bt.factor(1.0, 1e-6)
        Bt.solve()
        f, s_ = Bt.last_ms()
        f_ms.append(f)
        s_ms.append(s_)
    f_ms, s_ms = sorted(f_ms)[len(f_ms) // 2], sorted(s_ms)[len(s_ms) // 2]
    identical = bool(np.array_equal(Bt.get(0), sol1))
    inertia_ok = all((i.n_pos, i.n_neg, i.n_zero) == (fi.n_pos, fi.n_neg, 0) for i in info)
    pe, ue = Bt.stored_entries()
    Bt.close()
    P.close_device()
    P.close()
    gb = lambda by, ms: B * by / (ms * 1e-3) / 1e9
    out = {
        "instances": B, "kernels": "k_batch_factor + k_batch_solve (csrc/batch.cuh)",
        "arithmetic": "tensor (fused Schur updates)" if arithmetic else "reference",
        "factor_ms": f_ms, "solve_ms": s_ms,
        "factor": {"achieved": gb(fac_bytes, f_ms), "frac": gb(fac_bytes, f_ms) / peak},
        "solve": {"achieved": gb(sol_bytes, s_ms), "frac": gb(sol_bytes, s_ms) / peak},
        "factor_plus_solve": {
            "achieved": gb(fac_bytes + sol_bytes, f_ms + s_ms),
            "frac": gb(fac_bytes + sol_bytes, f_ms + s_ms) / peak},
        "with_min_degree_nnz_l": {
            "factor_plus_solve_frac": gb(fac_bytes_amd + sol_bytes_amd, f_ms + s_ms) / peak},
        "unit": "GB/s", "peak": peak,
        "algorithmic_bytes_per_instance": {"factor": fac_bytes, "solve": sol_bytes},
        "stored_doubles_per_instance": {"panels": pe, "updates": ue},
        "us_per_instance": (f_ms + s_ms) * 1e3 / B,
        "instance0_bit_identical_to_single_solve": identical,
        "inertia_of_every_instance_correct": inertia_ok,
        "timing": "median of 5 launches, CUDA events on the solver's stream; "
                  "inputs and factor exceed the L2 (no flush needed)",
    }
    return out


def run_sharded_leg(sb, dist, torch, rank, world, local_rank, warmup, steps,
                    N=20000, T=10.0):
    """BASELINE config 4: ONE cart-pole N = 20000 solve over all the GPUs of the
    run — the re-linearisation sweep split by time steps, every
    rank eliminating its own subtrees of the assembly tree, one small
    all-gather of the subtree roots, the top of the tree replicated, solution
    pieces gathered — against the same solve on one GPU. Strong scaling; the
    horizon is T = 10 s because with T = 5 s the reference algorithm leaves
    the Newton loop for feasibility restoration after 27 iterations
    (tests/test_gpu_configs.py)."""
    def timed(problem):
        problem.set_flush_l2(True)
        dist.barrier()
        torch.cuda.synchronize()
        problem.solve(max_iterations=warmup + steps, device=local_rank)
        tr = problem.trace()
        k, dt = steady_rate(tr, warmup, steps)
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return k, float(t.item())

    single = sb.Problem("cart_pole", N, T)
    k1, dt1 = timed(single)
    single.close()
    box = [sb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    P = sb.Problem("cart_pole", N, T)
    P.set_comm(rank, world, box[0])
    k2, dt2 = timed(P)
    cs = P.comm_stats()
    iters = max(len(P.trace()), 1)
    P.close()
    v1, v2 = k1 / dt1, k2 / dt2
    return {
        "workload": f"cart-pole direct transcription N={N}, T={T:g} s, ONE solve "
                    f"sharded over {world} GPUs (strong scaling)",
        "n_gpus": world, "value": v2, "unit": UNIT, "steps": k2,
        "ms_per_step": 1e3 * dt2 / k2,
        "single_gpu_value": v1, "single_gpu_ms_per_step": 1e3 * dt1 / k1,
        "speedup": v2 / v1, "efficiency": v2 / v1 / world,
        "allgather": {kind: {"calls_per_step": c["count"] / iters,
                             "bytes_per_call_per_rank": c["bytes"] / max(c["count"], 1),
                             "us_per_call": 1e3 * c["total_ms"] / max(c["count"], 1)}
                      for kind, c in cs.items()},
        "what": "per Newton iteration: one all-gather of the derivative rows "
                "(sweep split by tasks), one of the subtree roots per "
                "factorisation launch (update matrices, update vectors, inertia "
                "counts), one of the solution pieces per solve; NCCL channel "
                "set-up happens at slpb_comm_init, outside the timed window",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--horizon", type=int, default=5000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true",
                    help="experiments only: do not poll NVML during the run")
    ap.add_argument("--multistart", type=int, default=8,
                    help="also time B concurrent solves of the workload on each "
                         "GPU through slp::multistart (0: skip)")
    ap.add_argument("--batch", type=int, default=512,
                    help="instances of the batched many-instance LDLT leg "
                         "(slpb_batch_*: lane = instance; 0: skip)")
    ap.add_argument("--no-sharded-leg", action="store_true",
                    help="N > 1: skip the strong-scaling leg (ONE cart-pole "
                         "N=20000 solve sharded over the GPUs)")
    ap.add_argument("--shard", action="store_true",
                    help="N > 1: ONE solve sharded over the GPUs (derivative "
                         "sweep split over the ranks, one NCCL all-gather per "
                         "Newton iteration; strong scaling) instead of N "
                         "independent replicas")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.horizon
    workload = (f"cart-pole direct transcription N={N} (T=5 s, RK4, "
                f"n={5 * N + 4}, m_e={4 * N + 8}, m_i={4 * N + 2}), "
                "benchmark initial guess, default Options")
    cores = os.cpu_count()

    # ------------------------------------------------------------------ CPU arm
    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: at most 60 iterations (≈10 s) of the same workload.
        # The reference is single-threaded per solve; like our arm (one replica
        # per GPU) it gets one replica per requested GPU, each on its own core
        # (the reference's own multi-instance pattern, multistart.hpp:55).
        k = min(args.steps, 60)
        w = min(args.warmup, 3)
        n_rep = max(1, min(args.gpus, cores or 1))
        if n_rep == 1:
            results = [run_cpu(N, k, w)]
        else:
            import concurrent.futures as cf
            import multiprocessing as mp
            with cf.ProcessPoolExecutor(n_rep, mp_context=mp.get_context("spawn")) as ex:
                results = list(ex.map(run_cpu, [N] * n_rep, [k] * n_rep, [w] * n_rep,
                                      range(n_rep)))
        r = results[0]
        steps_done = min(x["steps"] for x in results)
        rate = sum(x["steps"] for x in results) / max(x["loop_s"] for x in results)
        # End to end like the GPU arm's e2e (one whole Problem::solve(), setup
        # amortised over ALL its iterations): the sample's setup time plus the
        # sample's steady rate carried over the iterations the reference
        # algorithm needs to reach its own exit status on this workload (the
        # committed golden; the sample's own count if there is none).
        full_iters, full_status = golden_iterations(N)
        if full_iters is None:
            full_iters = max(x["iters"] for x in results)
        setup_s = max(x["setup_s"] for x in results)
        e2e = n_rep * full_iters / (setup_s + full_iters * n_rep / rate)
        line = {
            "impl": "reference", "metric": METRIC, "value": rate,
            "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done, "warmup": w,
            "ms_per_step": 1e3 * n_rep / rate, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "host_cores_available": cores,
                       "parallelism": f"{n_rep} independent replica(s), one host "
                                      "core each"},
            "cpu_baseline": {
                "value": rate, "unit": UNIT, "cores": n_rep, "kind": "port",
                "sample": (f"iterations {w + 1}..{w + steps_done} of one solve of "
                           "the same workload per replica; oracle/ CPU restatement "
                           "of the reference IPM"
                           + (" running on the reference's OWN autodiff core "
                              "(oracle/_ref)" if r["backend"] == "reference"
                              else "")
                           + "; one thread per replica because the reference "
                             "path is single-threaded; each replica pinned to its "
                             f"own core (sched_setaffinity, core {r['core']} …)")},
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "what": (f"setup of the sample ({setup_s:.2f} s: problem "
                             "construction + autodiff setup) + its steady rate "
                             f"carried over {full_iters} iterations (what the "
                             "reference algorithm runs to its own exit status on "
                             "this workload, tests/golden/converge_cart_pole_"
                             f"{N}.npz"
                             + (f", status {full_status}" if full_status is not None else "")
                             + "): setup amortised over a whole solve, like the "
                             "GPU arm's e2e"),
                    "setup_s": setup_s, "iterations_projected": full_iters},
        }
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ GPU arm
    import torch
    import sleipnir_b200 as sb

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    peaks, peak_kind = measured_peaks()

    P = sb.Problem("cart_pole", N)
    # untimed: build + compile + first solve to warm the context / clocks
    P.solve(max_iterations=args.warmup, device=local_rank)
    P.close()

    shard = args.shard and world > 1

    def one_solve(flush, max_iterations, arithmetic=-1):
        Q = sb.Problem("cart_pole", N)
        Q.set_flush_l2(flush)
        Q.set_factor_arithmetic(arithmetic)
        if shard:
            # a ncclUniqueId serves ONE communicator: a fresh one per solve
            box = [sb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            Q.set_comm(rank, world, box[0])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        status = Q.solve(max_iterations=max_iterations, device=local_rank)
        torch.cuda.synchronize()
        return Q, time.perf_counter() - t0, status

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ClockSampler.enabled = not args.no_clock_sampler
    with ClockSampler(local_rank) as clk:
        # (1) headline `value`: L2 evicted before every iteration (256 MiB
        #     memset, excluded from the iteration timestamps)
        P, _, _ = one_solve(True, args.warmup + args.steps)
        tr = P.trace()
        k, dt = steady_rate(tr, args.warmup, args.steps)
        dt = max_over_ranks(dt)
        cnt, tim, sym = P.counters(), P.timers(), P.symbolic_stats()
        iters = len(tr)
        # (1b) the same steady-state leg with the LDLT in its tensor-core mode
        #     (SLPB_ARITH_TENSOR: fused Schur updates, DMMA on fronts ≥ 16)
        tensor = None
        if not shard:
            Pt, _, _ = one_solve(True, args.warmup + args.steps, sb.ARITH_TENSOR)
            trt = Pt.trace()
            kt, dtt = steady_rate(trt, args.warmup, args.steps)
            dtt = max_over_ranks(dtt)
            timt, cntt = Pt.timers(), Pt.counters()
            nl = max(timt["factor"]["launches"], 1)
            tensor = {"value": world * kt / dtt, "unit": UNIT, "steps": kt,
                      "factor_ms_per_launch": timt["factor"]["mean_ms"],
                      "factorizations_per_launch": cntt["factorizations_completed"] / nl}
            Pt.close()
        # (2) the solve as a user runs it: Problem::solve() with default Options
        #     (max_iterations = 5000) from host buffers TO CONVERGENCE, no
        #     flushes. Gives the end-to-end number and, from its first W+K
        #     iterations, the warm-L2 rate.
        P2, total_s, e2e_status = one_solve(False, 5000)
        tr2 = P2.trace()
        k2, dt2 = steady_rate(tr2, args.warmup, args.steps)
        dt2 = max_over_ranks(dt2)
        total_s = max_over_ranks(total_s)
        cnt2, phases = P2.counters(), P2.phase_seconds()
        # (3) the library's data-parallel axis: B independent solves in flight
        #     on this GPU (slp::multistart, one host thread + stream each).
        ms = None
        # (skipped when the ranks would not have a few host cores each for
        #  the starts' setup threads)
        if args.multistart > 1 and not shard and (cores or 1) // world >= 4:
            if world > 1:
                dist.barrier()
            ms = sb.multistart("cart_pole", N, [5.0] * args.multistart,
                               device=local_rank, n_vars=5 * N + 4)
            ms["wall_s"] = max_over_ranks(ms["wall_s"])
        sharded = None
        if world > 1 and not shard and not args.no_sharded_leg:
            sharded = run_sharded_leg(sb, dist, torch, rank, world, local_rank,
                                      args.warmup, min(args.steps, 100))
    replicas = 1 if shard else world   # independent solves running side by side
    rate = replicas * k / dt

    # ---- roofline -----------------------------------------------------------
    # Dominant kernel of ONE solve: k_factor_tree (supernodal LDLT). Algorithmic
    # bytes per numeric factorisation = 12·nnz(K) + 12·nnz(L) + 8·dim (SURVEY
    # §8d). Only factorisations that RAN TO COMPLETION are counted
    # (a speculated (0, 0) variant that meets a zero pivot is abandoned and moves
    # almost nothing); the figure is also given with the minimum-degree nnz(L)
    # (the reference's AMD order), so that the fill of the dissection order does
    # not inflate it.
    nnz_l_amd = amd_nnz_l(sb, N, local_rank)
    fac_bytes = 12 * sym["nnz_kkt"] + 12 * sym["nnz_l"] + 8 * sym["dim"]
    fac_bytes_amd = 12 * sym["nnz_kkt"] + 12 * nnz_l_amd + 8 * sym["dim"]
    # (the device timers sample one run in eight: mean_ms is over the sampled
    #  runs, `launches` counts all of them)
    n_launch = max(tim["factor"]["launches"], 1)
    fac_ms = tim["factor"]["mean_ms"]
    fac_per_launch = cnt["factorizations_completed"] / n_launch
    achieved = (fac_bytes * fac_per_launch) / (fac_ms * 1e-3) / 1e9 if fac_ms > 0 else 0.0
    achieved_amd = (fac_bytes_amd * fac_per_launch) / (fac_ms * 1e-3) / 1e9 if fac_ms > 0 else 0.0
    peak = float(peaks["hbm_gbs"])
    phase_ms = {k_: v["mean_ms"] for k_, v in tim.items()}
    per_step_ms = {k_: v["mean_ms"] * v["launches"] / iters for k_, v in tim.items()}
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    # k_solve_tree: the fused forward substitution rides in the factor launch,
    # so most solve launches are the backward half: 12·nnz(L) + 8·dim + 16·dim
    sol_bytes_full = 2 * 12 * sym["nnz_l"] + 8 * sym["dim"] + 4 * 8 * sym["dim"]
    sol_bytes_half = 12 * sym["nnz_l"] + 8 * sym["dim"] + 2 * 8 * sym["dim"]
    sol_ms = phase_ms["solve"]
    # autodiff sweep (second kernel of the step): algorithmic bytes of one full
    # re-linearisation = program stream + bindings + leaves + outputs
    ad_bytes = cnt["program_bytes"] + 8 * (5 * N + 4 + 8 * N + 10) + 8 * sym["nnz_kkt"]
    batch = None
    if world == 1 and args.batch > 0:
        batch = run_batch(sb, N, args.batch, local_rank, peak, fac_bytes,
                          sol_bytes_full, fac_bytes_amd,
                          2 * 12 * nnz_l_amd + 40 * sym["dim"])

    batch_tensor = None
    if world == 1 and args.batch > 0 and tensor is not None:
        batch_tensor = run_batch(sb, N, args.batch, local_rank, peak, fac_bytes,
                                 sol_bytes_full, fac_bytes_amd,
                                 2 * 12 * nnz_l_amd + 40 * sym["dim"],
                                 arithmetic=sb.ARITH_TENSOR)

    line = {
        "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": world,
        "steps": k, "warmup": args.warmup, "ms_per_step": 1e3 * dt / k,
        "higher_is_better": True, "scaling": "strong" if shard else "weak",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload,
            "parallelism": ("single GPU" if world == 1 else
                            f"ONE solve on {world} GPUs: derivative sweep sharded "
                            "over the ranks, one NCCL all-gather per Newton "
                            "iteration, factorisation and solve replicated"
                            if shard else
                            f"{world} independent replicas of the solve, one per "
                            "GPU, no data-path collective (DESIGN.md, Multi-GPU)"),
            "l2_policy": ("L2 flushed before every timed iteration (256 MiB "
                          "device memset > 126 MB L2, excluded from the iteration "
                          "timestamps); value_warm_l2 is the same solve without "
                          "flushes"),
            "value_warm_l2": replicas * k2 / dt2,
            "ordering": "nested dissection (level-set bisection)",
            "symbolic": sym,
            "per_step": {
                "factorizations": sum(r.factorizations for r in tr) / iters,
                "factor_launches": tim["factor"]["launches"] / iters,
                "solves": sum(r.solves for r in tr) / iters,
                "trial_points": sum(r.trials for r in tr) / iters},
            "device_ms_per_launch_group": phase_ms,
            "device_ms_per_step": per_step_ms,
            "solve_call_phases_s": phases,
        },
        "roofline": {
            "bound": "hbm",
            "kernel": "k_factor_tree (dependency-driven supernodal LDLT, one "
                      "launch per factorisation or speculated pair)",
            "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None,
            "traffic": traffic.get("k_factor_tree_dram_bytes_per_launch"),
            "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
            "algorithmic_bytes_per_factorization": fac_bytes,
            "factorizations_per_launch": fac_per_launch,
            "factorizations_per_launch_what": "completed numeric factorisations "
                "÷ launches (abandoned zero-pivot variants not counted)",
            "with_min_degree_nnz_l": {
                "nnz_l": nnz_l_amd, "algorithmic_bytes_per_factorization": fac_bytes_amd,
                "achieved": achieved_amd, "frac": achieved_amd / peak if peak else None},
            "ms_per_launch": fac_ms,
            "share_of_step_device_time": per_step_ms["factor"] / max(sum(per_step_ms.values()), 1e-12),
            "note": "latency-bound: the factor (≈7 MB) is L2-sized and the "
                    "assembly tree has 13 dependent levels; see DESIGN.md",
            "solve": {
                "kernel": "k_solve_tree (backward substitution; the forward half "
                          "is carried by the factor launch)",
                "algorithmic_bytes": sol_bytes_half, "ms": sol_ms,
                "achieved": sol_bytes_half / (sol_ms * 1e-3) / 1e9 if sol_ms > 0 else None,
                "frac": sol_bytes_half / (sol_ms * 1e-3) / 1e9 / peak if sol_ms > 0 else None,
                "traffic": traffic.get("k_solve_tree_dram_bytes_per_launch")},
            "batch": batch,
            "tensor_fronts": None if tensor is None else dict(tensor, **{
                "what": "the same steady-state leg with slpb_set_factor_arithmetic("
                        "SLPB_ARITH_TENSOR): every Schur-complement term one fused "
                        "multiply-add, rank-4 updates of frontal matrices of order "
                        ">= 16 on the FP64 tensor cores (mma.sync m8n8k4.f64, SASS "
                        "DMMA.8x8x4). Not the default: whole solves follow the "
                        "reference's regularisation decisions only with the "
                        "reference's separate rounding (DESIGN.md §3.3)",
                "achieved": (fac_bytes * tensor["factorizations_per_launch"]) /
                            (tensor["factor_ms_per_launch"] * 1e-3) / 1e9,
                "frac": (fac_bytes * tensor["factorizations_per_launch"]) /
                        (tensor["factor_ms_per_launch"] * 1e-3) / 1e9 / peak,
                "batch_factor_ms": None if batch_tensor is None else batch_tensor["factor_ms"],
                "batch_factor_frac": None if batch_tensor is None else batch_tensor["factor"]["frac"]}),
            "ad_sweep": {
                "traffic": traffic.get("k_ad_sweep_dram_bytes_per_launch"),
                "kernel": "k_ad_sweep (full re-linearisation)",
                "share_of_step_device_time": (per_step_ms["eval_full"] + per_step_ms["eval_values"]) /
                                             max(sum(per_step_ms.values()), 1e-12),
                "note": "as a kernel (derivative set + value set) k_ad_sweep takes about as "
                        "much of the step as k_factor_tree; it is bound by the dependency "
                        "depth of the expression graph (DESIGN.md §3.1), its algorithmic "
                        "bytes are negligible against the HBM peak",
                "algorithmic_bytes": ad_bytes,
                "ms": phase_ms["eval_full"],
                "achieved_gbs": ad_bytes / (phase_ms["eval_full"] * 1e-3) / 1e9 if phase_ms["eval_full"] > 0 else None},
        },
        "e2e": {
            "value": replicas * len(tr2) / total_s, "unit": UNIT,
            "what": "slp::Problem::solve() with default Options, run to its exit "
                    "status from HOST buffers: autodiff setup, tape upload, "
                    "compilation, symbolic analysis, every Newton iteration "
                    "(feasibility-restoration sub-solves included, with their "
                    "own setup) and the solution read-back; iterations ÷ wall "
                    "time of the call",
            "exit_status": sb.EXIT_STATUS[e2e_status],
            "iterations": len(tr2),
            "restoration_iterations": sum(1 for r in tr2 if r.type == 1),
            "h2d_bytes_per_step": cnt2["h2d_bytes"] / len(tr2),
            "d2h_bytes_per_step": cnt2["d2h_bytes"] / len(tr2),
            "solve_call_s": total_s},
        "sharded": sharded,
        "multistart": None if ms is None else {
            "value": world * sum(s_[2] for s_ in ms["starts"]) / ms["wall_s"],
            "unit": UNIT, "starts_per_gpu": args.multistart,
            "wall_s": ms["wall_s"],
            "exit_statuses": sorted({sb.EXIT_STATUS[s_[0]] for s_ in ms["starts"]}),
            "what": "slp::multistart (reference multistart.hpp:44-73): that many "
                    "independent Problem::solve() calls of the same workload in "
                    "flight per GPU, each on its own host thread, device handle "
                    "and CUDA stream, host setup included; total iterations ÷ "
                    "wall time of the call. One solve is latency-bound, so "
                    "concurrent solves overlap on the SMs"},
        "gpu_launches": int(cnt["kernel_launches"] * k / iters),
        "clocks": clk.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu(N, 12, 3)
        line["cpu_baseline"] = {
            "value": r["rate"], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": (f"iterations 4..{3 + r['steps']} of one solve of the same "
                       f"workload on the host ({cores} cores present, 1 used: the "
                       "reference path is single-threaded); oracle/ CPU "
                       "restatement of the reference IPM"
                       + (" on the reference's own autodiff core (oracle/_ref)"
                          if r["backend"] == "reference" else ""))}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
