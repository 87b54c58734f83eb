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

The N = 1 line also carries, each measured in the same run on the same box:
  ref_gpu : the UNMODIFIED reference fluid_solver_gpu (oracle/_ref/libref_gpu.so, compiled sm_100a) -- its own
            solve() at (Kd,Kp) = (15,20) and its stage sequence at the workload's K -- next to this repo's solve()
            and device-resident step at the SAME (Kd,Kp) ("the kernel to beat");
  c1, c2  : BASELINE configs[0] (256^2, K=20, 100 steps) and configs[1] (1024^2, K=40) with the current kernels.
The baselines that live under oracle/ (the unmodified fluid_solver_cpu and fluid_solver_gpu) are timed in CHILD
processes (`--leg ...`), so the process that runs this repo's solver never loads a checker library; the inputs
come from tools/canonical.py (a neutral generator).
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

import numpy as np  # noqa: E402,F401

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
    from tools import canonical as _c

    return _c.fields(n)


def run_leg(name, *extra, timeout=1200):
    """Run one baseline leg in a child process (`bench.py --leg NAME ...`) and parse its JSON line."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--leg", name] + [str(x) for x in extra],
                           capture_output=True, text=True, timeout=timeout)
        for ln in reversed(r.stdout.splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": ("rc %d: " % r.returncode) + (r.stderr.strip().splitlines() or ["no output"])[-1][:200]}
    except Exception as e:  # a baseline leg never takes the headline down with it
        return {"unavailable": str(e)[:200]}


# ------------------------------------------------------------------------------ baseline legs (child processes)
def cpu_reference_solver():
    """(callable(fields, kd, kp, nsteps) -> seconds, kind).  The unmodified fluid_solver_cpu when oracle/_ref
    travelled to this box, else the plain-C restatement of it."""
    from oracle import refs, sfo

    if refs.have_cpu():
        r = refs.ref_cpu()

        def run(f, kd, kp, nsteps=1):
            t = time.perf_counter()
            r.step_k(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, nsteps=nsteps)
            return time.perf_counter() - t

        return run, "reference"

    def run(f, kd, kp, nsteps=1):
        t = time.perf_counter()
        sfo.steps(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, smooth=False,
                  sem=sfo.SEM_CPU, nsteps=nsteps)
        return time.perf_counter() - t

    return run, "port"


def leg_cpu_baseline(argv):
    """--leg cpu_baseline N K STEPS: fluid_solver_cpu, one thread, STEPS steps of an N^2 grid with Kd=Kp=K."""
    n, k, steps = int(argv[0]), int(argv[1]), int(argv[2])
    run, kind = cpu_reference_solver()
    t = run(canonical(n), k, k, steps)
    print(json.dumps({"value": n * n * steps / t, "unit": UNIT, "cores": 1, "kind": kind, "ms_per_step": 1e3 * t / steps,
                      "sample": "%d step%s of a %dx%d grid, Kd=Kp=%d, fluid_solver_cpu (Gauss-Seidel, 1 thread; %d host cores present)"
                                % (steps, "" if steps == 1 else "s", n, n, k, os.cpu_count())}), flush=True)


def leg_ref_gpu(argv):
    """--leg ref_gpu N K STEPS [N K STEPS ...]: the unmodified fluid_solver_gpu on this box's GPU.  For each
    (N, K): its own solve() (Kd=15, Kp=20, src/fluid_solver_gpu.cu:222-258) and the same stage sequence with
    Kd=Kp=K (private-access shim, oracle/ref_gpu_shim.cu), one warm-up call, then STEPS steps between CUDA events
    (host<->device copies of solve() included: that IS the reference's step)."""
    from oracle import refs

    if not refs.have_gpu():
        print(json.dumps({"unavailable": "oracle/_ref/libref_gpu.so is not on this box (or no CUDA device)"}), flush=True)
        return
    g = refs.ref_gpu()
    out = {"what": "unmodified fluid_solver_gpu compiled -arch=sm_100a, default stream, pageable host grids; ms per step "
                   "from CUDA events around its solve() loop"}
    for a in range(0, len(argv), 3):
        n, k, steps = int(argv[a]), int(argv[a + 1]), int(argv[a + 2])
        f = canonical(n)
        args = (f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT)
        g.solve(*args, 1)
        ms_solve = g.solve(*args, steps)[3] / steps
        g.step_k(*args, k, k, True, 1)
        ms_k = g.step_k(*args, k, k, True, steps)[3] / steps
        out["n%d" % n] = {"solve_15_20_ms": ms_solve, "step_k_ms": ms_k, "k": k, "steps": steps}
    print(json.dumps(out), flush=True)


def run_reference_arm(args, workload):
    """The reference's CPU solver (single-threaded by design, src/fluid_solver_cpu.hpp:9) on a bounded
    sample: a 1024^2 grid with the workload's iteration counts; cell-steps/s is size-normalised.  One step at the
    workload's own grid is timed beside it (`same_grid`) when that takes under a minute and a half."""
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
    # the same solver on the workload's OWN grid, one step (4096^2, K=80: about half a minute); lets the
    # size-normalisation of the 1024^2 sample be checked against a same-grid figure
    n_w = workload["n"]
    same_grid = None
    est_s = float(n_w) * n_w / max(value, 1.0)
    if n_w != n_s and est_s <= 90.0:
        t_w = run(canonical(n_w), workload["kd"], workload["kp"])
        same_grid = {"grid": n_w, "steps": 1, "ms_per_step": 1e3 * t_w, "value": float(n_w) * n_w / t_w, "unit": UNIT}
    elif n_w != n_s:
        same_grid = {"grid": n_w, "skipped": "one step would take about %.0f s on one core" % est_s}
    sample = "%dx%d grid, Kd=Kp=%d, %d steps, fluid_solver_cpu (Gauss-Seidel, 1 thread)" % (n_s, n_s, workload["kd"], args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": workload["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(), "all_cores_replicas": replicas,
                         "same_grid_value": None if not same_grid else same_grid.get("value"),
                         "same_grid_ms_per_step": None if not same_grid else same_grid.get("ms_per_step")},
        "same_grid": same_grid,
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


def time_config(f2d, n, kd, kp, steps, e2e_steps, warmup=3, fields=None):
    """This repo at (n, kd, kp): device-resident ms/step (CUDA events) and solve() ms/call (pinned host grids, wall)."""
    fields = fields if fields is not None else canonical(n)
    with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, device=0) as s:
        s.upload(*fields[:3])
        s.set_sources(*fields[3:])
        s.step(DIFFUSION_RATE, VISCOSITY, DT, warmup)
        s.sync()
        ms = s.step_timed(DIFFUSION_RATE, VISCOSITY, DT, steps) / steps
        s.sync()
        pins = pinned_like(fields)
        hd, hu, hv, hsd, hsu, hsv = [p[1] for p in pins]
        for _ in range(2):
            s.solve(hd, hsd, DIFFUSION_RATE, hu, hv, hsu, hsv, VISCOSITY, DT)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            s.solve(hd, hsd, DIFFUSION_RATE, hu, hv, hsu, hsv, VISCOSITY, DT)
        e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    return ms, e2e_ms


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

    # ---- dominant kernel alone: the temporally blocked Jacobi pass, diffuse instance (the larger share of the
    #      step) and pressure instance, Kd / Kp sweeps each, CUDA events on the solver's stream
    reps = 5
    T = int(cfg.temporal_block)
    Td = int(cfg.temporal_block_diffuse)
    dif_ms = solver.bench_jacobi(True, kd, reps)
    dif_passes = sum(1 for _ in _passes(kd, Td)) * reps
    dif_gbs = 12.0 * cells * kd * reps / (dif_ms * 1e-3) / 1e9
    jac_ms = solver.bench_jacobi(False, kp, reps)
    passes = sum(1 for _ in _passes(kp, T)) * reps
    jac_gbs = 12.0 * cells * kp * reps / (jac_ms * 1e-3) / 1e9

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
    del pins, hd, hu, hv, hsd, hsu, hsv

    quick = bool(args.size or args.iters) or os.environ.get("F2D_BENCH_QUICK", "0") == "1"

    # ---- the unmodified reference GPU solver on this box (child process), and this repo at the SAME (Kd,Kp)
    ref_gpu, c1, c2 = None, None, None
    if not quick:
        rg = run_leg("ref_gpu", 256, 20, 100, 1024, 40, 5, n, kd, 2)
        if "unavailable" in rg:
            ref_gpu = rg
        else:
            r256, r1024, rn = rg["n256"], rg["n1024"], rg["n%d" % n]
            # this repo with the reference's own iteration counts (15, 20) on the workload grid
            o_ms, o_e2e = time_config(f2d, n, 15, 20, 10, 5, fields=fields)
            ref_gpu = {"grid": n, "what": rg["what"],
                       "ref_solve_15_20_ms": rn["solve_15_20_ms"], "ours_solve_15_20_ms": o_e2e, "ours_step_15_20_ms": o_ms,
                       "speedup_solve_15_20": rn["solve_15_20_ms"] / o_e2e, "speedup_step_15_20": rn["solve_15_20_ms"] / o_ms,
                       "ref_step_k%d_ms" % kd: rn["step_k_ms"], "ours_solve_k%d_ms" % kd: 1e3 * e2e_s / e2e_steps,
                       "ours_step_k%d_ms" % kd: ms / args.steps,
                       "speedup_solve_k%d" % kd: rn["step_k_ms"] / (1e3 * e2e_s / e2e_steps),
                       "speedup_step_k%d" % kd: rn["step_k_ms"] / (ms / args.steps)}
            # BASELINE configs[0]: 256^2, K=20, 100 steps (fluid_solver_cpu's own setting); configs[1]: 1024^2, K=40
            cpu1 = run_leg("cpu_baseline", 256, 20, 100)
            m1, e1 = time_config(f2d, 256, 20, 20, 100, 100)
            c1 = {"workload": "256x256, Kd=Kp=20, 100 steps (BASELINE configs[0])", "ms_per_step": m1,
                  "value": 256.0 * 256 / (m1 * 1e-3), "solve_ms": e1, "e2e_value": 256.0 * 256 / (e1 * 1e-3),
                  "ref_gpu_step_k_ms": r256["step_k_ms"], "ref_gpu_solve_15_20_ms": r256["solve_15_20_ms"],
                  "speedup_vs_ref_gpu_solve": r256["step_k_ms"] / e1, "speedup_vs_ref_gpu_step": r256["step_k_ms"] / m1,
                  "cpu_ms_per_step": cpu1.get("ms_per_step"), "cpu_value": cpu1.get("value"), "cpu_kind": cpu1.get("kind")}
            m2, e2 = time_config(f2d, 1024, 40, 40, 50, 20)
            c2 = {"workload": "1024x1024, Kd=Kp=40, 50 steps (BASELINE configs[1])", "ms_per_step": m2,
                  "value": 1024.0 * 1024 / (m2 * 1e-3), "solve_ms": e2, "e2e_value": 1024.0 * 1024 / (e2 * 1e-3),
                  "ref_gpu_step_k_ms": r1024["step_k_ms"], "ref_gpu_solve_15_20_ms": r1024["solve_15_20_ms"],
                  "speedup_vs_ref_gpu_solve": r1024["step_k_ms"] / e2, "speedup_vs_ref_gpu_step": r1024["step_k_ms"] / m2}

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
    #      values have to be divided by (the headline above is the 4096^2 roofline config, a different workload).
    #      Every N > 1 line also measures this base itself, in its own run (bench_multi.py).
    scaling_base = None
    if not quick and os.environ.get("F2D_BENCH_SCALING_BASE", "1") == "1":
        try:
            from bench_multi import single_gpu_base

            scaling_base = single_gpu_base(f2d, 16384, kd, kp, device=0)
        except Exception as e:  # e.g. not enough host memory on a small box: the headline does not depend on it
            scaling_base = {"unavailable": str(e)[:200]}

    # ---- CPU baseline on this box's host cores: bounded sample (one step of a 2048^2 grid), child process
    n_s = 2048 if kd >= 40 else 4096
    cpu = run_leg("cpu_baseline", n_s, kd, 1)

    bps = step_bytes(kd, kp)
    # real DRAM traffic per launch of the dominant kernel: from the committed ncu capture of this workload
    traffic, traffic_src, traffic_p = None, None, None
    for name in ("traffic_r02.json", "traffic_r01.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            tr = tj["k_jacobi_stream_diffuse"]
            if tr["grid"] == n and tr["temporal_block"] == Td:
                traffic, traffic_src = tr["dram_bytes_per_launch"], tr["source"]
                tp = tj.get("k_jacobi_stream_pressure", {})
                if tp.get("grid") == n and tp.get("temporal_block") == T:
                    traffic_p = tp["dram_bytes_per_launch"]
                break
        except (OSError, KeyError, ValueError):
            continue
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": workload["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload["name"], "grid": [n, n], "diffuse_iters": kd, "project_iters": kp,
                   "smooth": True, "dt": DT, "diffusion_rate": DIFFUSION_RATE, "viscosity": VISCOSITY,
                   "jacobi_mode": int(cfg.jacobi_mode), "temporal_block": T, "temporal_block_diffuse": Td, "divide_mode": int(cfg.divide_mode),
                   "cuda_graph": bool(cfg.use_graph),
                   "l2": "inputs larger than L2 (>= 13 fields x %.0f MiB)" % (cells * 4 / 2**20)},
        "roofline": {"bound": "hbm", "kernel": "k_jacobi_stream, diffuse instance (one field; the largest share of the step), %d sweeps per launch" % Td,
                     "achieved": dif_gbs, "peak": peak, "unit": "GB/s", "frac": dif_gbs / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 12.0 * cells * Td,
                     "peak_source": peak_src, "algorithmic_bytes_per_cell_sweep": 12,
                     "avg_launch_ms": dif_ms / dif_passes, "launches_timed": dif_passes,
                     "real_traffic_frac": None if not traffic else traffic / (dif_ms / dif_passes * 1e-3) / 1e9 / peak,
                     "pressure_achieved": jac_gbs, "pressure_frac": jac_gbs / peak, "pressure_avg_launch_ms": jac_ms / passes,
                     "pressure_traffic": traffic_p, "pressure_sweeps_per_launch": T,
                     "step_bytes_per_cell_step": bps, "step_achieved": bps * value / 1e9, "step_frac": bps * value / 1e9 / peak},
        "cpu_baseline": cpu if "unavailable" in cpu else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(6 * 4 * cells),
                "d2h_bytes_per_step": int(3 * 4 * cells), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "api": "FluidSolverB200.solve -> f2d_solve_host (pinned host grids; uploads, step parts and downloads overlapped)",
                "ref_gpu_step_same_k_ms": None if not ref_gpu else ref_gpu.get("ref_step_k%d_ms" % kd),
                "speedup_vs_ref_gpu_same_k": None if not ref_gpu else ref_gpu.get("speedup_solve_k%d" % kd)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "cpu_exact_mode": exact_mode,
        "scaling_base": scaling_base,
        "ref_gpu": ref_gpu,
        "c1": c1,
        "c2": c2,
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
    ap.add_argument("--leg", nargs=argparse.REMAINDER, help="internal: run one baseline leg in this (child) process")
    args = ap.parse_args()
    if args.leg:
        {"cpu_baseline": leg_cpu_baseline, "ref_gpu": leg_ref_gpu}[args.leg[0]](args.leg[1:])
        return
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
