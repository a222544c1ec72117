#!/usr/bin/env python
"""bench.py -- ODE-solve integrand evals/s of the batched expectation integrand on B200.

A "step" is one pass of the hot path over one batch of synthetic input:
    b200e_reduce_samples over the rank's sample set (h -> Tsit5 solve -> g -> sum) followed, for
    N > 1, by the all-reduce of the 2*nout partial sums + 4 counters (NCCL via torch.distributed).
Default workload = BASELINE.json configs[1]: the README 2-state linear ODE, MonteCarlo(10^7) on a
shared host-generated sample set (seeded, block-keyed; tests/sampling_util.py).  Each rank owns its
own 10^7-sample blocks of the stream ("scaling": "weak", no data-path collective besides the final
all-reduce of O(10) numbers).  Next to `value` every line carries
  * "strong": whole-solve wall times at a FIXED total (BASELINE configs[3] lorenz63 MonteCarlo(10^9), and configs[1]
    at its quoted 10^7 where the latency floor shows) split over the N ranks -- north_star's strong-scaling claim;
  * "in_library" (N > 1): the same legs through ONE process that owns all N GPUs (n_devices = N in b200e_model:
    ncclCommInitAll + one grouped ncclAllReduce inside the library) -- what a single-process caller such as the
    Julia shim gets;
  * "e2e_pageable": e2e from an ordinary (pageable) host array, what Integrals / Cubature hand f!(dx, x, p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

--impl reference times the CPU restatement of the reference path (oracle/, OpenMP over all host
cores -- the decomposition EnsembleThreads uses) on a bounded sample of the same workload: Julia is
not in this image, so the reference itself cannot run.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests.sampling_util import BLOCK, sample_block  # noqa: E402

WORKLOADS = {
    # name: (model, samples per rank per step, description)
    "linear2_mc1e7": ("linear2", 10_000_000, "README 2-state linear ODE du=A*u, MonteCarlo(10^7), Tsit5 abstol=1e-6 reltol=1e-3"),
    "lorenz63_mc": ("lorenz63", 16_000_000, "Lorenz-63, 3 uncertain params + 3 uncertain u0, MonteCarlo block of 1.6e7 (C4 streams 10^9 in such blocks)"),
    "lotka_volterra_mc": ("lotka_volterra", 8_000_000, "Lotka-Volterra 4 uncertain params, MonteCarlo(8e6)"),
    "oscillator_noise_mc": ("oscillator_noise", 2_000_000, "KKL process-noise damped oscillator, Euler-Maruyama dt=2^-10, MonteCarlo(2e6)"),
    "gbm4_lamba_mc": ("gbm4_lamba", 2_000_000, "test/processnoise.jl:9-15: du = u dt + u dW, 8-term KKL noise, adaptive LambaEM abstol=reltol=1e-3, MonteCarlo(2e6)"),
}
F_RHS = {"linear2": 6, "lotka_volterra": 6, "lorenz63": 8, "decay1": 1, "oscillator_noise": 7, "gbm4": 0, "gbm4_lamba": 0}


def oracle_problem(model: str, **kw):
    """The CPU restatement of a bench model (test infrastructure, used by the cpu_baseline / reference legs only)."""
    import oracle
    if model == "gbm4_lamba":
        return oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=1e-3, reltol=1e-3, **kw)
    return oracle.OracleProblem.builtin(model, **kw)


def algorithmic_flops(model: str, spec, counters: dict, npts: int) -> float:
    """F = nf*F_rhs + n_attempt*(70n+12) + npts*(12n+10) + npts*F_obs (FMA = 2)."""
    n = spec.nstate
    if spec.solver == 0:
        att = counters["naccept"] + counters["nreject"]
        return counters["nf"] * F_RHS[model] + att * (70 * n + 12) + npts * (12 * n + 10) + npts * spec.nout
    if spec.solver == 2:
        # LambaEM: per attempted step two drift + two diffusion evaluations, 16n for K, u', utilde, Ed, En, the
        # scaled residual and its square sum, 3*n_kkl for W(t+dt) (each sine counted as 1, like each pow) and the
        # 12 of the controller
        att = counters["naccept"] + counters["nreject"]
        return att * (4 * F_RHS[model] + 16 * n + 3 * spec.n_kkl + 12) + npts * (12 * n + 10) + npts * spec.nout
    # Euler-Maruyama: per step 2*n_kkl (dW from the table) + F_f + F_g + 4n
    return counters["naccept"] * (2 * spec.n_kkl + F_RHS[model] + 4 * n) + npts * spec.nout


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.005)  # a few samples per 20-step region without contending for the driver lock

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_samples(model: str, n: int, seed: int, first_block: int, out: np.ndarray):
    """Fill out[(n, nx)] with blocks first_block.. of the seeded stream."""
    got, b = 0, first_block
    while got < n:
        m = min(BLOCK, n - got)
        out[got:got + m] = sample_block(model, seed, b)[:m]
        got += m
        b += 1


def host_cores() -> int:
    """Threads the CPU arm uses: every core this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which must not throttle the reference arm, so the count is passed explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_restatement_rate(model: str, n_sample: int, seed: int = 7, reps: int = 1, x=None):
    """Times the oracle (CPU restatement of the reference path) on all host cores.  Returns
    (evals/s, cores, n, seconds)."""
    orc = oracle_problem(model)
    cores = host_cores()
    if x is None:
        x = np.empty((n_sample, orc.nx))
        make_samples(model, n_sample, seed, 0, x)
    x = x[:n_sample]
    orc.reduce_samples(x[: min(n_sample, 1 << 16)], nthreads=cores)  # warm-up (page in, spin threads)
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.reduce_samples(x, nthreads=cores)
    dt = time.perf_counter() - t0
    return n_sample * reps / dt, cores, n_sample, dt


KOOPMAN_CASES = {
    # name: (model, lb, ub, ODE tol, ireltol=iabstol, maxevals) -- "wall time per Expectation solve" (BASELINE metric)
    "koopman_readme_linear2": ("linear2", [1.0, 0.0], [10.0, 6.0], None, 1e-3, 1_000_000),       # configs[0], README.md:57-61
    "koopman_lotka_volterra": ("lotka_volterra", [1.0, 0.5, 2.0, 0.5], [2.0, 1.5, 4.0, 1.5], 1e-8, 2e-5, 3_000_000),  # configs[2]
}


def time_koopman_solves(arm: str, reps: int = 3):
    """Wall time of whole Koopman Expectation solves: the restated hcubature_v driver on the host, the
    batched integrand on the GPU (arm 'b200', through b200e_eval_nodes) or on the host cores (arm 'cpu',
    oracle).  Both arms see the same node batches."""
    import scimlexpectations_jl_b200 as sme
    out = {}
    for name, (model, lb, ub, otol, itol, maxevals) in KOOPMAN_CASES.items():
        kw = {} if otol is None else {"abstol": otol, "reltol": otol}
        if arm == "b200":
            h = sme.Handle(sme.models.builtin(model, **kw))
            f = lambda x: h.eval_nodes(x, weight_by_pdf=True)  # noqa: E731
        else:
            import oracle
            orc = oracle.OracleProblem.builtin(model, **kw)
            nth = host_cores()
            f = lambda x: orc.eval_nodes(x, weight_by_pdf=True, nthreads=nth)[0]  # noqa: E731
        nout = 2 if model == "linear2" else 1
        best, res = None, None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = sme.hcubature_v(f, lb, ub, fdim=nout, reltol=itol, abstol=itol, maxevals=maxevals)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out[name] = {"wall_ms": best * 1e3, "u": res.u.tolist(), "resid": res.resid.tolist(), "nevals": res.nevals,
                     "rounds": res.nrounds, "max_batch": max(res.batch_sizes), "converged": res.converged,
                     "ode_tol": otol or "default 1e-6/1e-3", "ireltol": itol,
                     "driver": "hcubature_v restated on the host (cubature.py), integrand per round through "
                               + ("b200e_eval_nodes" if arm == "b200" else "the oracle (OpenMP)")}
        if arm == "b200":
            # the same solve as ONE C call: heap in the library, Genz-Malik rule on the device
            bestd = None
            for _ in range(reps):
                t0 = time.perf_counter()
                u, resid, st = h.koopman_solve(lb, ub, reltol=itol, abstol=itol, maxevals=maxevals)
                dt = time.perf_counter() - t0
                bestd = dt if bestd is None else min(bestd, dt)
            out[name + "_device_rule"] = {"wall_ms": bestd * 1e3, "u": u.tolist(), "resid": resid.tolist(),
                                          "nevals": st["nevals"], "rounds": st["nrounds"], "max_batch": st["max_batch"],
                                          "converged": bool(st["converged"]), "kernel_ms": st["kernel_ms"],
                                          "driver": "b200e_koopman_solve"}
            h.close()
    return out


def core_config(args, model, n_full, desc, nx):
    """The keys that name the workload: identical in both arms (the driver compares them)."""
    return {"workload": args.workload, "model": model, "description": desc,
            "samples_per_gpu_per_step": int(n_full),
            "sample_set": f"PCG64(SeedSequence([{args.seed}, block])), blocks of 2^20",
            "l2": f"input {n_full * nx * 8 / 1e6:.0f} MB per GPU > 126 MB L2" if n_full * nx * 8 > 126e6 else "input smaller than L2"}


def gpu_numa_cpus(index: int):
    """CPUs of the NUMA node the GPU hangs off (None when the box has one node or sysfs does not say)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        return node, (cpus or None)
    except Exception:
        return None, None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    model, n_full, desc = WORKLOADS[args.workload]
    orc = oracle_problem(model)
    cores = host_cores()
    # bounded sample per step: ~2 s of CPU work, calibrated on a small block
    x_cal = np.empty((1 << 17, orc.nx))
    make_samples(model, len(x_cal), 7, 0, x_cal)
    orc.reduce_samples(x_cal, nthreads=cores)
    t0 = time.perf_counter()
    orc.reduce_samples(x_cal, nthreads=cores)
    rate = len(x_cal) / (time.perf_counter() - t0)
    n = int(min(n_full, max(1 << 17, rate * 2.0)))
    x = np.empty((n, orc.nx))
    make_samples(model, n, args.seed, 0, x)
    for _ in range(args.warmup):
        orc.reduce_samples(x, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s, s2, c = orc.reduce_samples(x, nthreads=cores)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "ODE-solve integrand evals/sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": core_config(args, model, n_full, desc, orc.nx),
        "reference_note": "CPU restatement of the reference's EnsembleThreads path (Julia is not installed in this "
                          "image, the reference cannot run); each step is a bounded sample of the workload",
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"first {n} samples of the workload's sample set per step, OpenMP over {cores} threads"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_solves:
        line["expectation_solves"] = time_koopman_solves("cpu")
    emit_json_line(line)
    return 0


STRONG_CASES = (
    # name, model, FIXED total samples of the whole solve, warm-up samples per rank
    ("lorenz63_mc1e9", "lorenz63", 10**9),      # BASELINE configs[3]
    ("linear2_mc1e7", "linear2", 10**7),        # BASELINE configs[1] at its quoted size: the latency floor
)


def strong_scaling_legs(sme, args, world, rank, local_rank, barrier):
    """north_star: '>85 % strong-scaling efficiency from 1 to 8 GPUs'.  The TOTAL sample count of a solve is fixed;
    rank r of R draws and solves indices [r n/R, (r+1) n/R) of the counter-based stream inside its kernel
    (b200e_reduce_generated), one all-reduce of the 2 nout + 4 numbers finishes the solve.  Wall time of the whole
    solve: barrier + synchronize on both sides, max over ranks, best of `reps`.  The expectation and the integer
    RHS-evaluation count must not depend on R (sample <-> index mapping is independent of the sharding)."""
    import torch
    import torch.distributed as dist
    if args.no_strong:
        return None
    out = {}
    for name, model, total in STRONG_CASES:
        lo, hi = total * rank // world, total * (rank + 1) // world
        with sme.Handle(sme.models.builtin(model, devices=[local_rank])) as hs:
            nout = hs.spec.nout
            red = torch.zeros(2 * nout + 4, dtype=torch.float64, device="cuda")
            stage = torch.zeros(2 * nout + 4, dtype=torch.float64).pin_memory()

            def solve():
                s, s2, c = hs.reduce_generated(args.seed, hi - lo, first_index=lo)
                a = stage.numpy()
                a[:nout], a[nout:2 * nout] = s, s2
                a[2 * nout:] = (c["nf"], c["naccept"], c["nreject"], c["nfail"])
                red.copy_(stage, non_blocking=True)
                if world > 1:
                    dist.all_reduce(red)
                return red.cpu().numpy()

            hs.reduce_generated(1, min(hi - lo, 10**6))   # module load, buffers
            solve()
            reps = 3 if total >= 10**8 else 20
            times = []
            for _ in range(reps):
                barrier()
                t0 = time.perf_counter()
                r = solve()
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                times.append(float(dt.item()))
            kms = hs.last_stats()["kernel_ms"]
        best = min(times)
        out[name] = {"total_samples": total, "n_gpus": world, "wall_s": best, "median_wall_s": float(np.median(times)),
                     "evals_per_s": total / best, "kernel_ms_rank0": kms, "reps": reps,
                     "expectation": (r[:nout] / total).tolist(), "nf": int(r[2 * nout]), "nfail": int(r[2 * nout + 3]),
                     "api": "b200e_reduce_generated per rank + one all-reduce (torch.distributed, NCCL)"}
    return out


def in_library_legs(sme, args, world, rank, model, n, nx, x_rank0):
    """The drop-in mode of a single-process caller (the reference consumes ONE ensemblealg in ONE process,
    src/expectation.jl:191, :290): one handle with n_devices = N; the library shards the samples in contiguous index
    ranges over its devices and finishes with ncclCommInitAll's communicator -- one grouped ncclAllReduce of the
    2 nout + 4 numbers (b200_expect.cu exchange_results).  Rank 0 drives all N GPUs while the other ranks wait on a
    CPU (gloo) barrier with their GPUs idle."""
    import torch
    import torch.distributed as dist
    if world == 1 or args.no_inlib:
        return None
    cpu_group = dist.new_group(backend="gloo")
    res = None
    torch.cuda.synchronize()
    dist.barrier(group=cpu_group)
    if rank == 0:
        res = {"n_devices": world, "process": "one (rank 0), ncclCommInitAll + grouped ncclAllReduce inside libb200expect"}
        devs = list(range(world))

        def timeit(fn, steps, warmup=3):
            for _ in range(warmup):
                fn()
            t0 = time.perf_counter()
            for _ in range(steps):
                out = fn()
            return (time.perf_counter() - t0) / steps, out

        with sme.Handle(sme.models.builtin(model, devices=devs)) as hm:
            nout = hm.spec.nout
            # (1) the public MonteCarlo call, rand(d) in the kernel: N x n samples per call (weak, like e2e_device_sampling)
            dt, (s, s2, c) = timeit(lambda: hm.reduce_generated(args.seed, world * n), args.steps)
            res["e2e_device_sampling"] = {"value": world * n / dt, "unit": "evals/s", "ms_per_step": dt * 1e3,
                                          "expectation": (s / (world * n)).tolist(), "nf": c["nf"]}
            # (2) host-fed: one pinned buffer of N x n samples, sharded by the library over its devices
            xall = hm.pinned((world * n, nx))
            for r in range(world):
                nb = (n + BLOCK - 1) // BLOCK
                make_samples(model, n, args.seed, r * nb, xall.array[r * n:(r + 1) * n])
            e_steps = max(3, args.steps // 4)
            dt, (s, s2, c) = timeit(lambda: hm.reduce_samples(xall), e_steps)
            st = hm.last_stats()
            res["e2e"] = {"value": world * n / dt, "unit": "evals/s", "ms_per_step": dt * 1e3,
                          "h2d_bytes_per_step": int(st["h2d_bytes"]), "expectation": (s / (world * n)).tolist(),
                          "nf": c["nf"], "api": "b200e_reduce_samples(host pinned x, n_devices = N) -> host sums"}
            xall.free()
        if not args.no_strong:
            res["strong"] = {}
            for name, smodel, total in STRONG_CASES:
                with sme.Handle(sme.models.builtin(smodel, devices=devs)) as hs:
                    hs.reduce_generated(1, world * 10**6)
                    reps = 3 if total >= 10**8 else 20
                    times = []
                    for _ in range(reps):
                        t0 = time.perf_counter()
                        s, s2, c = hs.reduce_generated(args.seed, total)
                        times.append(time.perf_counter() - t0)
                    res["strong"][name] = {"total_samples": total, "wall_s": min(times), "median_wall_s": float(np.median(times)),
                                           "evals_per_s": total / min(times), "expectation": (s / total).tolist(),
                                           "nf": c["nf"], "nfail": c["nfail"]}
    dist.barrier(group=cpu_group)
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    import scimlexpectations_jl_b200 as sme

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; this path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    model, n, desc = WORKLOADS[args.workload]
    if args.samples:
        n = args.samples
    spec = sme.models.builtin(model, devices=[local_rank], refill_threshold=args.refill,
                              block_size=args.block, blocks_per_sm=args.blocks_per_sm,
                              fp_mode=1 if args.fp32 else 0)
    h = sme.Handle(spec)
    nx, nout = spec.nx, spec.nout

    # rank r owns blocks [r*nb, (r+1)*nb) of the seeded stream: same sets the CPU oracle sees in tests
    nb = (n + BLOCK - 1) // BLOCK
    # host-fed legs: the rank's pinned buffer is allocated and first touched on the NUMA node of its GPU
    numa_node, numa_cpus = gpu_numa_cpus(local_rank)
    old_aff = os.sched_getaffinity(0)
    if numa_cpus:
        os.sched_setaffinity(0, numa_cpus)
    xp = h.pinned((n, nx))
    make_samples(model, n, args.seed, rank * nb, xp.array)
    xpage = np.array(xp.array, copy=True)   # an ordinary (pageable) array: what f!(dx, x, p) receives from Cubature
    if numa_cpus:
        os.sched_setaffinity(0, old_aff)
    xd = h.device_array((n, nx)).upload(xp.array)

    # the path's one exchange step: all-reduce of the 2*nout sums + 4 counters (counters < 2^53 as doubles)
    red = torch.zeros(2 * nout + 4, dtype=torch.float64, device="cuda")
    # two pinned staging buffers, alternated, so that a step never rewrites one an async copy may still read
    red_hs = [torch.zeros(2 * nout + 4, dtype=torch.float64).pin_memory() for _ in range(2)]
    red_nps = [t.numpy() for t in red_hs]
    red_evs = [torch.cuda.Event() for _ in range(2)]
    xstate = {"i": 0}

    def exchange(s, s2, c):
        i = xstate["i"] = xstate["i"] ^ 1
        red_evs[i].synchronize()  # the copy that last read this buffer (two steps ago) is long done
        a = red_nps[i]
        a[:nout] = s
        a[nout:2 * nout] = s2
        a[2 * nout:] = (c["nf"], c["naccept"], c["nreject"], c["nfail"])
        red.copy_(red_hs[i], non_blocking=True)
        red_evs[i].record()
        dist.all_reduce(red)

    def step_resident():
        s, s2, c = h.reduce_samples(xd, npts=n)
        st = h.last_stats()
        if world > 1:
            exchange(s, s2, c)
        return s, s2, c, st

    def step_e2e():
        s, s2, c = h.reduce_samples(xp)  # pinned host buffer: H2D inside the call, sums come back D2H
        st = h.last_stats()
        if world > 1:
            exchange(s, s2, c)
            torch.cuda.synchronize()
        return s, s2, c, st

    def step_pageable():
        s, s2, c = h.reduce_samples(xpage)  # pageable host memory: the library stages it (host-fed path)
        st = h.last_stats()
        if world > 1:
            exchange(s, s2, c)
            torch.cuda.synchronize()
        return s, s2, c, st

    def step_generated():
        # solve(exprob, MonteCarlo(n)) as the user calls it: rand(d) is drawn inside the kernel from the
        # counter-based stream (seed, index); rank r owns indices [r*n, (r+1)*n)
        s, s2, c = h.reduce_generated(args.seed, n, first_index=rank * n)
        st = h.last_stats()
        if world > 1:
            exchange(s, s2, c)
            torch.cuda.synchronize()
        return s, s2, c, st

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        kernel_ms, total_ms, launches = [], [], 0
        t0 = time.perf_counter()
        for _ in range(steps):
            s, s2, c, st = fn()
            kernel_ms.append(st["kernel_ms"])
            total_ms.append(st["total_ms"])
            launches += st["launches"]
        torch.cuda.synchronize()
        t_local = time.perf_counter() - t0
        barrier()
        t = torch.tensor([t_local], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), kernel_ms, total_ms, launches, (s, s2, c, st)

    def timed_resident(steps, warmup):
        """The timed region of `value`: K steps, each = one launch over the resident samples + (N > 1) its own
        all-reduce of the 2 nout + 4 results.  The library's split call keeps the kernel of step k+1 queued behind
        step k (as a streamed MonteCarlo(10^9) does, block after block), so the device does not wait for the host
        between steps and the exchange of step k runs on the host and over NVLink while step k+1 computes; every
        step still finishes with its own exchange, the last one inside the timed region."""
        kernel_ms, total_ms, launches = [], [], 0
        prev, st = None, None

        def retire():
            nonlocal prev, launches, st
            prev = h.reduce_samples_end()
            st = h.last_stats()
            kernel_ms.append(st["kernel_ms"])
            total_ms.append(st["total_ms"])
            launches += st["launches"]
            if world > 1:
                exchange(*prev)

        def run_steps(k):
            h.reduce_samples_begin(xd, npts=n)
            for _ in range(k - 1):
                h.reduce_samples_begin(xd, npts=n)   # step j+1 queues behind step j (two pending reductions per handle)
                retire()                              # result of step j, then its exchange while step j+1 runs
            retire()

        if warmup > 0:
            run_steps(warmup)                         # the same pipelined pattern: both reduce slots exist before t0
        torch.cuda.synchronize()
        barrier()
        del kernel_ms[:], total_ms[:]
        launches = 0
        t0 = time.perf_counter()
        run_steps(steps)
        torch.cuda.synchronize()
        t_local = time.perf_counter() - t0
        barrier()
        t = torch.tensor([t_local], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), kernel_ms, total_ms, launches, (prev[0], prev[1], prev[2], st)

    sampler = ClockSampler(local_rank)
    sampler.start()
    T, kernel_ms, total_ms, launches, (s, s2, c, st) = timed_resident(args.steps, args.warmup)
    clocks = sampler.stop()
    e_steps = max(3, args.steps // 4)
    Te, _, e_total_ms, _, (se, s2e, ce, ste) = timed(step_e2e, e_steps, 3)
    # the ceiling of every host-fed leg: plain cudaMemcpy of the rank's pinned buffer, all ranks at once
    def step_h2d_only():
        xd.upload(xp.array)
        return None, None, None, {"kernel_ms": 0.0, "total_ms": 0.0, "launches": 0}
    Th, _, _, _, _ = timed(step_h2d_only, e_steps, 2)
    h2d_ceiling_gbs = world * n * nx * 8 * e_steps / Th / 1e9
    Tp, _, _, _, (sp_, _, cp_, stp) = timed(step_pageable, e_steps, 3)
    Tg, g_kernel_ms, _, _, (sg, s2g, cg, stg) = timed(step_generated, args.steps, 3)
    strong = strong_scaling_legs(sme, args, world, rank, local_rank, barrier)
    in_library = in_library_legs(sme, args, world, rank, model, n, nx, xp.array if world > 1 else None)

    # per-rank device times and clocks (outside the timed regions): the job moves at the pace of its slowest GPU
    per_rank = None
    if world > 1:
        mine = torch.tensor([float(np.mean(kernel_ms)), float(np.mean(total_ms)), float(clocks.get("sm_mhz") or 0.0)],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernel_ms": [round(float(t[0]), 4) for t in allr], "device_ms_per_step": [round(float(t[1]), 4) for t in allr],
                    "sm_mhz": [float(t[2]) for t in allr]}

    value = world * n * args.steps / T
    e2e_value = world * n * e_steps / Te
    k_ms = float(np.mean(kernel_ms))
    flops = algorithmic_flops(model, spec, c, n)
    peak = sme.measure_fma_peak(local_rank, args.fp32)
    achieved = flops / (k_ms * 1e-3) / 1e12
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_path)).get("hbm_gbs") if os.path.exists(peaks_path) else 6650.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)

    if rank == 0:
        mean = s / n
        line = {
            "metric": "ODE-solve integrand evals/sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": T / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.fp32 else "f64", "data": "synthetic",
            "config": core_config(args, model, n, desc, nx),
            "launch": {"grid": st["grid"], "block": st["block"], "regs": st["regs_per_thread"],
                       "local_bytes_per_thread": st["local_bytes_per_thread"],
                       "refill_threshold": args.refill or "auto", "parallelism": f"samples sharded over {world} GPU(s)",
                       "pinned_buffer_numa_node": numa_node},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(ste["h2d_bytes"]),
                    "d2h_bytes_per_step": int(ste["d2h_bytes"]), "ms_per_step": Te / e_steps * 1e3,
                    "api": "b200e_reduce_samples(host pinned x) -> host sums",
                    # what the host can feed at all: the same buffers through plain cudaMemcpy, every rank at once
                    "h2d_ceiling_gbs": h2d_ceiling_gbs, "h2d_achieved_gbs": world * n * nx * 8 * e_steps / Te / 1e9},
            # not the contract's e2e (its inputs never exist on the host): the public MonteCarlo call with the
            # sampler inside the kernel, the configuration a user of solve(exprob, MonteCarlo(n)) runs
            # the same call from pageable host memory (a numpy array; Julia's Array handed to f!(dx, x, p))
            "e2e_pageable": {"value": world * n * e_steps / Tp, "unit": "evals/s", "h2d_bytes_per_step": int(stp["h2d_bytes"]),
                             "d2h_bytes_per_step": int(stp["d2h_bytes"]), "ms_per_step": Tp / e_steps * 1e3,
                             "slowdown_vs_pinned": Tp / Te,
                             "api": "b200e_reduce_samples(host pageable x) -> host sums"},
            "strong": strong,
            "e2e_device_sampling": {"value": world * n * args.steps / Tg, "unit": "evals/s", "h2d_bytes_per_step": 0,
                                    "d2h_bytes_per_step": int(stg["d2h_bytes"]), "ms_per_step": Tg / args.steps * 1e3,
                                    "kernel_ms": float(np.mean(g_kernel_ms)),
                                    "expectation": (sg / n).tolist(),
                                    "api": "b200e_reduce_generated(seed, first_index, n) -> host sums "
                                           "(Philox4x32-10 stream drawn in registers; b200e_sample_host reproduces it bit for bit)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64" if not args.fp32 else "fp32", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         # the instantiation the library launches for this batch (dynamic work distribution
                         # whenever the batch holds more than one point per lane, b200_expect.cu use_dynamic)
                         "kernel": "b200e_reduce_dyn" if n > st["grid"] * st["block"] else "b200e_reduce",
                         "kernel_ms": k_ms,
                         "flops_per_launch": flops,
                         "peak_source": "b200e_measure_fma_peak (FMA-chain microbenchmark run in this process; "
                                        "MEASURED_PEAKS.json has no FP64 entry)",
                         "hbm": {"achieved": n * nx * 8 / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s"}},
            "counters": c, "expectation": mean.tolist(),
            "device_ms_per_step": float(np.mean(total_ms)),
        }
        if per_rank is not None:
            line["per_rank"] = per_rank
        if in_library is not None:
            line["in_library"] = in_library
        if world == 1 and not args.no_solves:
            line["expectation_solves"] = time_koopman_solves("b200")
        if world == 1 and not args.no_cpu:
            # ~10-20 s of CPU work on the box's host cores
            r0, cores, _, _ = cpu_restatement_rate(model, 1 << 18, x=xp.array)
            ncpu = int(min(n, max(1 << 18, r0 * 12.0)))
            reps = max(1, int(round(r0 * 12.0 / ncpu)))
            rate, cores, ncpu, secs = cpu_restatement_rate(model, ncpu, reps=reps, x=xp.array)
            line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                                    "sample": f"first {ncpu} samples of the same sample set x {reps} passes, {secs:.1f} s, "
                                              "oracle/ (C restatement of the reference path, OpenMP over all host cores)"}
        emit_json_line(line)
    xd.free()
    h.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def keep_stdout_for_the_json_line():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on
    the first collective when the box exports NCCL_DEBUG), so everything else that goes to file descriptor 1 is sent
    to stderr and the JSON line alone is written to the original stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_json_line(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="linear2_mc1e7", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU per step")
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--refill", type=int, default=0)
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-solves", action="store_true", help="skip the Koopman whole-solve timings")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-total (strong scaling) whole-solve legs")
    ap.add_argument("--no-inlib", action="store_true", help="skip the single-process n_devices = N legs (N > 1)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    keep_stdout_for_the_json_line()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
