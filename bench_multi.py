"""Multi-GPU leg of bench.py: one process per GPU (torchrun), row slabs, NCCL halo exchange.
value = N^2 * steps / max-over-ranks(device time); strong scaling on the 16384^2 grid."""
import json
import os
import time

import numpy as np

from bench import DIFFUSION_RATE, DT, METRIC, UNIT, VISCOSITY, ClockSampler, measured_peaks, step_bytes


def canonical_rows(n, r0, r1):
    """Rows [r0, r1) of the canonical fields, generated locally (no full-grid host arrays)."""
    from tools import canonical as _c

    return list(_c.rows(n, r0, r1))


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, before any host buffer exists: the pinned slabs
    are then first-touched on that node, and solve()'s 36 B per cell of PCIe traffic does not cross the socket link.
    Host-side placement only (the reference has no notion of it); silently skipped where sysfs or the affinity call
    are not available."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return {"bound": False, "numa_node": node}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": str(e)[:120]}


def single_gpu_base(f2d, n, kd, kp, device=0, steps=5):
    """The N > 1 workload (n x n, Kd, Kp) on ONE GPU, device-resident: the denominator of strong-scaling efficiency.
    Measured inside the run that reports the efficiency (rank 0, after the slab solvers are gone)."""
    fb = canonical_rows(n, 0, n)
    with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, device=device) as sb:
        sb.upload(*fb[:3])
        sb.set_sources(*fb[3:])
        del fb
        sb.step(DIFFUSION_RATE, VISCOSITY, DT, 3)
        sb.sync()
        b_ms = sb.step_timed(DIFFUSION_RATE, VISCOSITY, DT, steps) / steps
        sb.sync()
    return {"workload": "%dx%d grid, Kd=Kp=%d on ONE GPU (the workload of the N > 1 runs), device-resident" % (n, n, kd),
            "value": float(n) * n / (b_ms * 1e-3), "unit": UNIT, "ms_per_step": b_ms, "steps": steps,
            "note": "strong-scaling efficiency of an N-GPU line = its value / (N * this value)"}


def run_multi_gpu(args, workload):
    import torch
    import torch.distributed as dist

    import fluid2d_b200 as f2d
    from fluid2d_b200 import slab as slabmod

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if os.environ.get("F2D_BENCH_NUMA_BIND", "1") == "1" else {"bound": False}
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, kd, kp = workload["n"], workload["kd"], workload["kp"]
    halo = int(os.environ.get("F2D_HALO", "32"))
    cfl = 8
    sl = slabmod.partition(n, world, halo, rank)
    transport = os.environ.get("F2D_TRANSPORT", "auto")
    tdev = torch.device("cuda", local_rank)
    uid = slabmod.broadcast_unique_id(dist, rank, device=tdev) if transport == "nccl" else None
    solver = slabmod.make_slab_solver(sl, n, uid, cfl_cells=cfl, device=local_rank, transport=transport, dist=dist,
                                      torch_device=tdev, diffuse_iters=kd, project_iters=kp)
    cfg = solver.config()
    d, u, v, sd, su, sv = canonical_rows(n, sl.row_offset, sl.row_offset + sl.rows)
    solver.upload(d, u, v)
    solver.set_sources(sd, su, sv)

    def barrier():
        solver.sync()
        dist.barrier()
        torch.cuda.synchronize()

    solver.step(DIFFUSION_RATE, VISCOSITY, DT, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0, x0, b0 = solver.launch_count(), slabmod.comm_exchanges(solver), slabmod.comm_bytes(solver)
    ms = solver.step_timed(DIFFUSION_RATE, VISCOSITY, DT, args.steps)
    barrier()
    launches, xch = solver.launch_count() - l0, slabmod.comm_exchanges(solver) - x0
    halo_stats = {"bytes_per_neighbour_per_step": (slabmod.comm_bytes(solver) - b0) / max(1, args.steps)}
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    fast = os.environ.get("F2D_BENCH_FAST", "0") == "1"  # weak-scaling sweeps: skip the e2e and CPU legs
    if fast:
        solver.close()
        if rank == 0:
            cells = float(n) * n
            value = cells * args.steps / (ms_max * 1e-3)
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                              "scaling": "weak", "dtype": "f32", "data": "synthetic",
                              "config": {"workload": workload["name"], "grid": [n, n], "diffuse_iters": kd, "project_iters": kp,
                                         "transport": getattr(solver, "transport", transport), "halo": halo},
                              "gpu_launches": int(launches), "exchanges_per_step": xch / max(1, args.steps), "clocks": clocks}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    # end to end: every rank uploads its slab from pinned host memory, steps once, downloads it
    pins = [torch.from_numpy(a).pin_memory() for a in (d, u, v, sd, su, sv)]
    hd, hu, hv, hsd, hsu, hsv = [p.numpy() for p in pins]
    e2e_steps = 3
    for _ in range(1):
        solver.solve(hd, hsd, DIFFUSION_RATE, hu, hv, hsu, hsv, VISCOSITY, DT)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.solve(hd, hsd, DIFFUSION_RATE, hu, hv, hsu, hsv, VISCOSITY, DT)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    solver.close()

    if rank == 0:
        cells = float(n) * n
        value = cells * args.steps / (ms_max * 1e-3)
        peak, peak_src = measured_peaks()
        bps = step_bytes(kd, kp)
        from bench import run_leg

        # strong-scaling base, same run, same box: the SAME grid and K on one GPU (the other ranks wait in the barrier)
        base = None
        if os.environ.get("F2D_BENCH_SCALING_BASE", "1") == "1":
            try:
                base = single_gpu_base(f2d, n, kd, kp, device=local_rank)
            except Exception as e:
                base = {"unavailable": str(e)[:200]}
        speedup = efficiency = None
        if base and "value" in base:
            speedup = value / base["value"]
            efficiency = speedup / world
        # NVLink traffic of the halo exchanges: bytes this rank pushes to ONE neighbour per step, and the rate if the
        # whole step's push were spread over the step time (the exchanges are latency-, not bandwidth-bound)
        push_bytes = halo_stats["bytes_per_neighbour_per_step"]
        nvlink_peak = 900.0  # GB/s per direction per GPU, nominal (B200_PROFILING.md; measured peer copy 770)
        cpu = run_leg("cpu_baseline", 2048, kd, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": workload["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload["name"], "grid": [n, n], "diffuse_iters": kd, "project_iters": kp,
                       "parallelism": "row slabs x%d, halo %d rows, %s" % (world, halo, "direct NVLink peer stores + epoch-flag handshake in the step graph" if getattr(solver, "transport", transport) == "p2p" else "NCCL send/recv in the step graph"),
                       "transport": getattr(solver, "transport", transport),
                       "temporal_block": int(cfg.temporal_block), "temporal_block_diffuse": int(cfg.temporal_block_diffuse), "jacobi_mode": int(cfg.jacobi_mode),
                       "divide_mode": int(cfg.divide_mode), "cfl_cells": cfl,
                       "scaling_base_ms_per_step": None if not base else base.get("ms_per_step"),
                       "scaling_base_value": None if not base else base.get("value"),
                       "speedup_vs_1gpu_same_grid": speedup, "efficiency": efficiency,
                       "l2": "inputs larger than L2 (slab fields of %.0f MiB)" % (sl.rows * n * 4 / 2**20)},
            "roofline": {"bound": "hbm", "kernel": "whole step, algorithmic bytes (SURVEY 8d)", "achieved": bps * value / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": bps * value / 1e9 / (peak * world), "traffic": None,
                         "peak_source": peak_src + " x n_gpus",
                         "halo_exchanges_per_step": xch / max(1, args.steps), "halo_rows": halo,
                         "halo_bytes_per_neighbour_per_step": push_bytes,
                         "nvlink_gbs_avg_over_step": push_bytes / (ms_max / args.steps * 1e-3) / 1e9,
                         "nvlink_frac_of_900": push_bytes / (ms_max / args.steps * 1e-3) / 1e9 / nvlink_peak,
                         "nvlink_time_at_900_ms": push_bytes / (nvlink_peak * 1e9) * 1e3},
            "cpu_baseline": cpu if "unavailable" in cpu else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cells * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(6 * 4 * cells),
                    "d2h_bytes_per_step": int(3 * 4 * cells), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": "FluidSolverB200.solve per slab (pinned host slabs, halo rows included)",
                    "host_numa_binding_rank0": numa},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "scaling_base": base,
            "speedup": speedup,
            "efficiency": efficiency,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
