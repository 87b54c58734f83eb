#!/usr/bin/env python
"""bench.py -- cell-steps/s of the stable-fluids step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full solve() step (add_source, diffuse, advect, project) over the whole grid.
  N = 1 : 4096 x 4096, Kd = Kp = 80   (BASELINE.json configs[2], the single-GPU roofline config)
  N > 1 : 16384 x 16384, Kd = Kp = 80, row slabs over N GPUs (configs[3]); strong scaling
Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs in HBM, CUDA events
on the solver's stream, max over ranks); `e2e` is the same metric through the reference interface
`solve()` with pinned HOST grids, all host<->device copies inside the timed region.
`--impl reference` times the reference's own CPU solver (oracle/_ref/libref_cpu.so, or the oracle
port when it is absent) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DT, DIFFUSION_RATE, VISCOSITY = 0.02, 0.5, 1e-6  # src/app.cpp:8,33-34
METRIC, UNIT = "cell_steps_per_sec", "cell-steps/s"


def step_bytes(kd, kp, smooth=True):
    """Streaming-ideal HBM bytes per cell-step."""
    return (148 if smooth else 140) + 36 * kd + 24 * kp


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def canonical(n):
    from oracle import sfo

    return sfo.canonical_fields(n)


def cpu_reference_solver():
    """(callable(fields, kd, kp) -> seconds per step, kind, cores)."""
    from oracle import refs, sfo

    if refs.have_cpu():
        r = refs.ref_cpu()

        def run(f, kd, kp):
            t = time.perf_counter()
            r.step_k(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, nsteps=1)
            return time.perf_counter() - t

        return run, "reference"

    def run(f, kd, kp):
        t = time.perf_counter()
        sfo.steps(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, smooth=False,
                  sem=sfo.SEM_CPU, nsteps=1)
        return time.perf_counter() - t

    return run, "port"


def run_reference_arm(args, workload):
    """The reference's CPU solver (single-threaded by design, src/fluid_solver_cpu.hpp:9) on a bounded
    sample: a 1024^2 grid with the workload's iteration counts; cell-steps/s is size-normalised."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = 1024
    run, kind = cpu_reference_solver()
    f = canonical(n_s)
    for _ in range(args.warmup):
        run(f, workload["kd"], workload["kp"])
    t = sum(run(f, workload["kd"], workload["kp"]) for _ in range(args.steps))
    value = n_s * n_s * args.steps / t
    # fluid_solver_cpu is single-threaded by design; the most the host can do with it is one independent replica
    # per core.  Reported beside the line's value (which stays the solver as shipped), never instead of it.
    ncores = os.cpu_count() or 1
    times = [0.0] * ncores

    def replica(k):
        times[k] = run(f, workload["kd"], workload["kp"])

    ths = [threading.Thread(target=replica, args=(k,)) for k in range(ncores)]
    t0 = time.perf_counter()
    [th.start() for th in ths]
    [th.join() for th in ths]
    wall = time.perf_counter() - t0
    replicas = {"value": ncores * n_s * n_s / wall, "unit": UNIT, "cores": ncores,
                "what": "%d independent single-threaded replicas of the same step, one per host core, aggregate" % ncores}
    sample = "%dx%d grid, Kd=Kp=%d, %d steps, fluid_solver_cpu (Gauss-Seidel, 1 thread)" % (n_s, n_s, workload["kd"], args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": workload["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(), "all_cores_replicas": replicas},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pinned_like(arrs):
    """Pinned host copies (torch is only plumbing here: page-locked allocation)."""
    import torch

    out = []
    for a in arrs:
        t = torch.from_numpy(a.copy()).pin_memory()
        out.append((t, t.numpy()))
    return out


def run_single_gpu(args, workload):
    import fluid2d_b200 as f2d

    n, kd, kp = workload["n"], workload["kd"], workload["kp"]
    if f2d.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (libf2d has no CPU fallback)")
    fields = canonical(n)
    d, u, v, sd, su, sv = fields
    peak, peak_src = measured_peaks()
    cells = float(n) * n

    solver = f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, device=0)
    cfg = solver.config()
    solver.upload(d, u, v)
    solver.set_sources(sd, su, sv)
    # ---- device-resident throughput: W warm-up steps, K timed steps between CUDA events
    if args.warmup:
        solver.step(DIFFUSION_RATE, VISCOSITY, DT, args.warmup)
    solver.sync()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    l0 = solver.launch_count()
    ms = solver.step_timed(DIFFUSION_RATE, VISCOSITY, DT, args.steps)
    launches = solver.launch_count() - l0
    solver.sync()
    clocks = sampler.stop()
    value = cells * args.steps / (ms * 1e-3)

    # ---- dominant kernel alone: the temporally blocked Jacobi pass (pressure solve, Kp sweeps)
    reps = 5
    jac_ms = solver.bench_jacobi(False, kp, reps)
    T = int(cfg.temporal_block)
    passes = sum(1 for _ in _passes(kp, T)) * reps
    jac_cell_iters = cells * kp * reps
    jac_gbs = 12.0 * jac_cell_iters / (jac_ms * 1e-3) / 1e9
    dif_ms = solver.bench_jacobi(True, kd, reps)
    dif_gbs = 12.0 * cells * kd * reps / (dif_ms * 1e-3) / 1e9
    Td = int(cfg.temporal_block_diffuse)

    # ---- end to end through the reference interface: solve() on pinned host grids
    pins = pinned_like(fields)
    hd, hu, hv, hsd, hsu, hsv = [p[1] for p in pins]
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        solver.solve(hd, hsd, DIFFUSION_RATE, hu, hv, hsu, hsv, VISCOSITY, DT)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.solve(hd, hsd, DIFFUSION_RATE, hu, hv, hsu, hsv, VISCOSITY, DT)
    e2e_s = time.perf_counter() - t0
    e2e_value = cells * e2e_steps / e2e_s
    solver.close()

    # ---- the fluid_solver_cpu-compatible mode (F2D_SEM_CPU, bit-identical to the reference arm's solver) on the
    #      reference arm's own sample (1024^2, the workload's K): what "the same bits, on the GPU" costs
    n_x = 1024
    fx = canonical(n_x)
    with f2d.FluidSolverB200.cpu_compatible(n_x, n_x, iters=kd, device=0) as sx:
        sx.upload(*fx[:3])
        sx.set_sources(*fx[3:])
        sx.step(DIFFUSION_RATE, VISCOSITY, DT, 2)
        sx.sync()
        x_ms = sx.step_timed(DIFFUSION_RATE, VISCOSITY, DT, 5) / 5
        sx.sync()
    exact_mode = {"value": n_x * n_x / (x_ms * 1e-3), "unit": UNIT, "ms_per_step": x_ms,
                  "workload": "%dx%d grid, K=%d Gauss-Seidel sweeps, fluid_solver_cpu arithmetic (F2D_SEM_CPU), device-resident" % (n_x, n_x, kd),
                  "parity": "bit-identical to fluid_solver_cpu::solve (tests/test_gpu_cpu_semantics.py)"}

    # ---- strong-scaling base: the N > 1 runs use the 16384^2 grid (configs[3]); its single-GPU time is what their
    #      values have to be divided by (the headline above is the 4096^2 roofline config, a different workload)
    scaling_base = None
    if not args.size and not args.iters and os.environ.get("F2D_BENCH_SCALING_BASE", "1") == "1":
        try:
            from bench_multi import canonical_rows

            nb = 16384
            fb = canonical_rows(nb, 0, nb)
            with f2d.FluidSolverB200(nb, nb, diffuse_iters=kd, project_iters=kp, device=0) as sb:
                sb.upload(*fb[:3])
                sb.set_sources(*fb[3:])
                del fb
                sb.step(DIFFUSION_RATE, VISCOSITY, DT, 3)
                sb.sync()
                b_ms = sb.step_timed(DIFFUSION_RATE, VISCOSITY, DT, 5) / 5
                sb.sync()
            scaling_base = {"workload": "16384x16384 grid, Kd=Kp=%d on ONE GPU (the workload of the N > 1 runs)" % kd,
                            "value": float(nb) * nb / (b_ms * 1e-3), "unit": UNIT, "ms_per_step": b_ms, "steps": 5,
                            "note": "strong-scaling efficiency of an N-GPU line = its value / (N * this value)"}
        except Exception as e:  # e.g. not enough host memory on a small box: the headline does not depend on it
            scaling_base = {"unavailable": str(e)[:200]}

    # ---- CPU baseline on this box's host cores: bounded sample (one step of a 2048^2 grid)
    run, kind = cpu_reference_solver()
    n_s = 2048 if kd >= 40 else 4096
    cpu_t = run(canonical(n_s), kd, kp)
    cpu_value = n_s * n_s / cpu_t

    bps = step_bytes(kd, kp)
    # real DRAM traffic per launch of the dominant kernel: from the committed ncu capture of this workload
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))["k_jacobi_stream_pressure"]
        if tr["grid"] == n and tr["temporal_block"] == T:
            traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
    except (OSError, KeyError, ValueError):
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": workload["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload["name"], "grid": [n, n], "diffuse_iters": kd, "project_iters": kp,
                   "smooth": True, "dt": DT, "diffusion_rate": DIFFUSION_RATE, "viscosity": VISCOSITY,
                   "jacobi_mode": int(cfg.jacobi_mode), "temporal_block": T, "temporal_block_diffuse": int(cfg.temporal_block_diffuse), "divide_mode": int(cfg.divide_mode),
                   "cuda_graph": bool(cfg.use_graph),
                   "l2": "inputs larger than L2 (>= 13 fields x %.0f MiB)" % (cells * 4 / 2**20)},
        "roofline": {"bound": "hbm", "kernel": "k_jacobi_stream (pressure relaxation, %d sweeps per launch)" % T,
                     "achieved": jac_gbs, "peak": peak, "unit": "GB/s", "frac": jac_gbs / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 12.0 * cells * T,
                     "peak_source": peak_src, "algorithmic_bytes_per_cell_sweep": 12,
                     "avg_launch_ms": jac_ms / passes, "launches_timed": passes,
                     "diffuse_kernel": {"achieved": dif_gbs, "frac": dif_gbs / peak, "sweeps_per_launch": Td},
                     "step": {"bytes_per_cell_step": bps, "achieved": bps * value / 1e9, "frac": bps * value / 1e9 / peak}},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": "1 step of a %dx%d grid, Kd=Kp=%d, fluid_solver_cpu (Gauss-Seidel, 1 thread; %d host cores present)"
                                   % (n_s, n_s, kd, os.cpu_count())},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(6 * 4 * cells),
                "d2h_bytes_per_step": int(3 * 4 * cells), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "api": "FluidSolverB200.solve -> f2d_solve_host (pinned host grids; uploads, step parts and downloads overlapped)"},
        "cpu_exact_mode": exact_mode,
        "scaling_base": scaling_base,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def _passes(k, T):
    left = k
    while left > 0:
        t = T
        while t > left:
            t >>= 1
        yield t
        left -= t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="override the grid size (debug)")
    ap.add_argument("--iters", type=int, default=0, help="override Kd = Kp (debug)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.gpus <= 1:
        workload = {"n": 4096, "kd": 80, "kp": 80, "scaling": "strong",
                    "name": "4096x4096 grid, 80 Jacobi iters (Kd=Kp=80), 1xB200 (BASELINE configs[2])"}
    else:
        workload = {"n": 16384, "kd": 80, "kp": 80, "scaling": "strong",
                    "name": "16384x16384 grid, 80 Jacobi iters, row slabs over %d B200 (BASELINE configs[3])" % args.gpus}
    if args.size:
        workload["n"] = args.size
        workload["name"] = "%dx%d grid (override)" % (args.size, args.size)
    if args.iters:
        workload["kd"] = workload["kp"] = args.iters
        workload["name"] += ", Kd=Kp=%d (override)" % args.iters

    if args.impl == "reference":
        run_reference_arm(args, workload)
        return
    if args.gpus <= 1:
        run_single_gpu(args, workload)
    else:
        from bench_multi import run_multi_gpu

        run_multi_gpu(args, workload)


if __name__ == "__main__":
    main()
