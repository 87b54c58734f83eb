"""Weak-scaling sweep of BASELINE configs[4]: 2^27 cells per GPU, Kd = Kp in {20, 40, 80, 200}, device-resident.
  1 GPU :  python tools/weak_sweep.py                      (11584^2 on one GPU)
  8 GPUs:  torchrun --nproc-per-node 8 tools/weak_sweep.py (32768^2 on row slabs)
One process per GPU loops over K (one solver per K), so the process start-up is paid once.  Prints one JSON line per K."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import canonical  # noqa: E402

DT, RATE, VISC = 0.02, 0.5, 1e-6
KS = [int(x) for x in os.environ.get("F2D_WEAK_KS", "20,40,80,200").split(",")]
STEPS, WARMUP = 5, 3


def main():
    import fluid2d_b200 as f2d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        n = int(os.environ.get("F2D_WEAK_N", "11584"))
        f = canonical.rows(n, 0, n)
        for k in KS:
            with f2d.FluidSolverB200(n, n, diffuse_iters=k, project_iters=k) as s:
                s.upload(*f[:3])
                s.set_sources(*f[3:])
                s.step(RATE, VISC, DT, WARMUP)
                s.sync()
                ms = s.step_timed(RATE, VISC, DT, STEPS) / STEPS
                s.sync()
            print(json.dumps({"n_gpus": 1, "grid": n, "k": k, "ms_per_step": ms, "cell_steps_per_s": float(n) * n / (ms * 1e-3)}), flush=True)
        return
    import torch
    import torch.distributed as dist

    from fluid2d_b200 import slab as slabmod

    rank, local_rank = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = int(os.environ.get("F2D_WEAK_N", "32768"))
    halo = 32
    sl = slabmod.partition(n, world, halo, rank)
    transport = os.environ.get("F2D_TRANSPORT", "p2p")
    tdev = torch.device("cuda", local_rank)
    f = canonical.rows(n, sl.row_offset, sl.row_offset + sl.rows)
    for k in KS:
        uid = slabmod.broadcast_unique_id(dist, rank, device=tdev) if transport == "nccl" else None
        s = slabmod.make_slab_solver(sl, n, uid, cfl_cells=8, device=local_rank, transport=transport, dist=dist, torch_device=tdev,
                                     diffuse_iters=k, project_iters=k)
        s.upload(*f[:3])
        s.set_sources(*f[3:])
        s.step(RATE, VISC, DT, WARMUP)
        s.sync()
        dist.barrier()
        torch.cuda.synchronize()
        ms = s.step_timed(RATE, VISC, DT, STEPS)
        s.sync()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s.close()
        dist.barrier()
        if rank == 0:
            ms1 = float(t.item()) / STEPS
            print(json.dumps({"n_gpus": world, "grid": n, "k": k, "transport": transport, "ms_per_step": ms1,
                              "cell_steps_per_s": float(n) * n / (ms1 * 1e-3)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
