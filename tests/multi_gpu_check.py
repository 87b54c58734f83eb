"""Launched with torchrun on >= 2 GPUs: the row-slab path must reproduce the single-GPU result.
Deterministic fields (u, v) bit-for-bit; density within the scatter's summation-order tolerance.
Rank 0 also runs the single-GPU solver on the full grid as the reference of this self-consistency
check and compares it with the oracle at a small size."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import fluid2d_b200 as f2d
    from fluid2d_b200 import slab as slabmod
    from util import DIFFUSION_RATE, DT, VISCOSITY, err, rng_fields

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ok = True
    report = []
    cases = [(512, 15, 20, 16, 2, 8, True), (1024, 40, 40, 32, 2, 8, True), (768, 7, 9, 8, 3, 4, False), (2048, 80, 80, 32, 2, 8, True)]
    for n, kd, kp, halo, steps, T, graph in cases:
        fields = rng_fields(n, 5000 + n, vel_cells=4.0)  # same seed on every rank
        sl = slabmod.partition(n, world, halo, rank)
        transport = os.environ.get("F2D_TRANSPORT", "auto")
        tdev = torch.device("cuda", local_rank)
        uid = slabmod.broadcast_unique_id(dist, rank, device=tdev) if transport == "nccl" else None
        s = slabmod.make_slab_solver(sl, n, uid, cfl_cells=6, device=local_rank, transport=transport, dist=dist, torch_device=tdev,
                                     diffuse_iters=kd, project_iters=kp, temporal_block=T, divide_mode=f2d.DIV_F64, use_graph=graph)
        loc = [slabmod.take(sl, a) for a in fields]
        s.upload(*loc[:3])
        s.set_sources(*loc[3:])
        s.step(DIFFUSION_RATE, VISCOSITY, DT, steps)
        s.sync()
        out = s.download()
        xch = slabmod.comm_exchanges(s)
        s.close()
        # gather owned rows on rank 0
        glob = [np.zeros((n, n), np.float32) for _ in range(3)]
        for k in range(3):
            b, e = sl.local_own
            own = torch.from_numpy(np.ascontiguousarray(out[k][b:e])).cuda()
            parts = [torch.empty((slabmod.partition(n, world, halo, r).own_end - slabmod.partition(n, world, halo, r).own_begin, n),
                                 dtype=torch.float32, device="cuda") for r in range(world)]
            dist.all_gather(parts, own) if len({p.shape for p in parts}) == 1 else _gather_uneven(dist, parts, own, rank, world)
            if rank == 0:
                glob[k] = torch.cat(parts).cpu().numpy()
        if rank == 0:
            with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, temporal_block=T, divide_mode=f2d.DIV_F64,
                                     device=local_rank) as one:
                one.upload(*fields[:3])
                one.set_sources(*fields[3:])
                one.step(DIFFUSION_RATE, VISCOSITY, DT, steps)
                one.sync()
                ref = one.download()
            eu, ev, ed = err(glob[1], ref[1]), err(glob[2], ref[2]), err(glob[0], ref[0])
            good = eu["n_diff"] == 0 and ev["n_diff"] == 0 and ed["rel_l2"] <= 2e-6 * steps and ed["max_abs"] <= 2e-5 * steps * max(1.0, float(np.abs(ref[0]).max()))
            ok &= good
            report.append(dict(transport=transport, n=n, kd=kd, kp=kp, halo=halo, steps=steps, T=T, graph=graph, world=world, exchanges=xch,
                               u=eu, v=ev, d=ed, ok=bool(good)))
        dist.barrier()
    if rank == 0:
        print(json.dumps(report, indent=1))
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def _gather_uneven(dist, parts, own, rank, world):
    for r in range(world):
        if r == rank:
            parts[r].copy_(own)
        dist.broadcast(parts[r], src=r)


if __name__ == "__main__":
    main()
