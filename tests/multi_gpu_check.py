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
    transport = os.environ.get("F2D_TRANSPORT", "auto")
    tdev = torch.device("cuda", local_rank)

    def make(sl, n, kd, kp, T, graph, cfl):
        uid = slabmod.broadcast_unique_id(dist, rank, device=tdev) if transport == "nccl" else None
        return slabmod.make_slab_solver(sl, n, uid, cfl_cells=cfl, device=local_rank, transport=transport, dist=dist, torch_device=tdev,
                                        diffuse_iters=kd, project_iters=kp, temporal_block=T, divide_mode=f2d.DIV_F64, use_graph=graph)

    def gather_owned(sl, n, halo, out):
        glob = [None, None, None]
        for k in range(3):
            b, e = sl.local_own
            own = torch.from_numpy(np.ascontiguousarray(out[k][b:e])).cuda()
            parts = [torch.empty((slabmod.partition(n, world, halo, r).own_end - slabmod.partition(n, world, halo, r).own_begin, n),
                                 dtype=torch.float32, device="cuda") for r in range(world)]
            dist.all_gather(parts, own) if len({p.shape for p in parts}) == 1 else _gather_uneven(dist, parts, own, rank, world)
            if rank == 0:
                glob[k] = torch.cat(parts).cpu().numpy()
        return glob

    def any_rank_failed(failed):
        t = torch.tensor([1 if failed else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item()) == 1

    def single_gpu(n, kd, kp, T, steps, fields, via_solve):
        with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, temporal_block=T, divide_mode=f2d.DIV_F64,
                                 device=local_rank) as one:
            if via_solve:
                h = [a.copy() for a in fields]
                for _ in range(steps):
                    one.solve(h[0], h[3], DIFFUSION_RATE, h[1], h[2], h[4], h[5], VISCOSITY, DT)
                return h[:3]
            one.upload(*fields[:3])
            one.set_sources(*fields[3:])
            one.step(DIFFUSION_RATE, VISCOSITY, DT, steps)
            one.sync()
            return one.download()

    def judge(tag, glob, ref, steps, extra):
        eu, ev, ed = err(glob[1], ref[1]), err(glob[2], ref[2]), err(glob[0], ref[0])
        good = eu["n_diff"] == 0 and ev["n_diff"] == 0 and ed["rel_l2"] <= 2e-6 * steps and ed["max_abs"] <= 2e-5 * steps * max(1.0, float(np.abs(ref[0]).max()))
        report.append(dict(case=tag, transport=transport, world=world, u=eu, v=ev, d=ed, ok=bool(good), **extra))
        return good

    # ---- 1. device-resident stepping: slabs == one GPU
    cases = [(512, 15, 20, 16, 2, 8, True), (1024, 40, 40, 32, 2, 8, True), (768, 7, 9, 8, 3, 4, False), (2048, 80, 80, 32, 2, 8, True)]
    for n, kd, kp, halo, steps, T, graph in cases:
        fields = rng_fields(n, 5000 + n, vel_cells=4.0)  # same seed on every rank
        sl = slabmod.partition(n, world, halo, rank)
        s = make(sl, n, kd, kp, T, graph, 6)
        loc = [slabmod.take(sl, a) for a in fields]
        s.upload(*loc[:3])
        s.set_sources(*loc[3:])
        s.step(DIFFUSION_RATE, VISCOSITY, DT, steps)
        s.sync()
        out = s.download()
        xch = slabmod.comm_exchanges(s)
        s.close()
        glob = gather_owned(sl, n, halo, out)
        if rank == 0:
            ok &= judge("step", glob, single_gpu(n, kd, kp, T, steps, fields, False), steps,
                        dict(n=n, kd=kd, kp=kp, halo=halo, steps=steps, T=T, graph=graph, exchanges=xch))
            print(json.dumps(report[-1]), flush=True)
        dist.barrier()

    # ---- 2. solve() per slab (uploads, the four parts of the step with their exchanges, downloads overlapped) == one GPU.
    #         1024^2 slabs are >= 1 MiB and take the pipelined path; the host slabs carry stale halo rows between calls.
    #         912^2 on 4 ranks and 1280^2 on 8 put the edge slabs below and the interior slabs above the 1 MiB threshold
    #         of the pipelined path: every rank must still take the same path (same exchange schedule).
    for n, kd, kp, halo, steps, T, graph in [(1024, 15, 20, 32, 2, 8, True), (1280, 24, 16, 32, 2, 8, False), (912, 15, 20, 32, 2, 8, True)]:
        fields = rng_fields(n, 6000 + n, vel_cells=4.0)
        sl = slabmod.partition(n, world, halo, rank)
        s = make(sl, n, kd, kp, T, graph, 0)
        h = [slabmod.take(sl, a) for a in fields]
        failed, msg = False, ""
        for it in range(steps):
            try:
                s.solve(h[0], h[3], DIFFUSION_RATE, h[1], h[2], h[4], h[5], VISCOSITY, DT)
            except f2d.F2DError as e:
                failed, msg = True, "rank %d, call %d: %s" % (rank, it, e)
                print("SOLVE FAILED", msg, flush=True)
                break
        s.close()
        if any_rank_failed(failed):
            if rank == 0:
                report.append(dict(case="solve", transport=transport, world=world, n=n, graph=graph, ok=False, error=msg or "on another rank"))
                ok = False
            dist.barrier()
            continue
        glob = gather_owned(sl, n, halo, h[:3])
        if rank == 0:
            ok &= judge("solve", glob, single_gpu(n, kd, kp, T, steps, fields, True), steps,
                        dict(n=n, kd=kd, kp=kp, halo=halo, steps=steps, T=T, graph=graph))
            print(json.dumps(report[-1]), flush=True)
        dist.barrier()

    # ---- 3. the displacement bound (CFL) is verified on the device: never a silently wrong state.
    #   a) ~20 rows per step with the default bound (halo - 1 = 31): equal to one GPU;
    #   b) ~28 rows per step with a promised bound of 6 rows: the promise is broken -> f2d_sync fails on some rank;
    #   c) ~45 rows per step: beyond what a 32-row halo can serve -> f2d_sync fails.
    n, kd, kp, halo, T = 1024, 8, 8, 32, 8
    for tag, cells, cfl, expect_error in (("cfl_ok_20_of_31", 20.0, 0, False), ("cfl_promise_6_broken", 28.0, 6, True),
                                          ("cfl_beyond_halo", 45.0, 0, True)):
        fields = list(rng_fields(n, 7000, vel_cells=cells))
        # every cell moves 0.9 .. 1.0 x `cells` rows (v > 0 everywhere), so the rows next to a slab edge certainly do
        fields[2] = (np.float32(0.95 * cells / (n * DT)) + np.float32(0.05) * fields[2]).astype(np.float32)
        fields[5] = np.zeros_like(fields[5])
        sl = slabmod.partition(n, world, halo, rank)
        s = make(sl, n, kd, kp, T, True, cfl)
        loc = [slabmod.take(sl, a) for a in fields]
        s.upload(*loc[:3])
        s.set_sources(*loc[3:])
        failed = False
        try:
            s.step(DIFFUSION_RATE, VISCOSITY, DT, 1)
            s.sync()
        except f2d.F2DError as e:
            failed = True
            msg = str(e)
        out = s.download()
        s.close()
        failed_any = any_rank_failed(failed)
        if expect_error:
            good = failed_any
            if rank == 0:
                report.append(dict(case=tag, transport=transport, world=world, error_reported=failed_any, ok=bool(good)))
                ok &= good
        else:
            glob = gather_owned(sl, n, halo, out)
            if rank == 0:
                good = (not failed_any) and judge(tag, glob, single_gpu(n, kd, kp, T, 1, fields, False), 1, dict(n=n, cells=cells))
                ok &= good
        dist.barrier()
    # ---- 4. (F2D_CHECK_BIG=n) a published configuration against the UNMODIFIED reference: n x n canonical fields,
    #         Kd = Kp = 80, one step on row slabs (fp64 divide), owned rows gathered on rank 0 and compared with
    #         fluid_solver_gpu's own stage sequence at the same K (oracle/_ref/libref_gpu.so, run live on rank 0's GPU).
    big = int(os.environ.get("F2D_CHECK_BIG", "0"))
    if big:
        from tools import canonical as canon

        n, kd, kp, halo, T = big, 80, 80, 32, 8
        sl = slabmod.partition(n, world, halo, rank)
        loc = list(canon.rows(n, sl.row_offset, sl.row_offset + sl.rows))
        s = make(sl, n, kd, kp, T, True, 0)
        s.upload(*loc[:3])
        s.set_sources(*loc[3:])
        s.step(DIFFUSION_RATE, VISCOSITY, DT, 1)
        s.sync()
        out = s.download()
        s.close()
        del loc
        glob = gather_owned(sl, n, halo, out)
        del out
        torch.cuda.empty_cache()
        if rank == 0:
            from oracle import refs

            if refs.have_gpu():
                f = list(canon.rows(n, 0, n))
                ref = refs.ref_gpu().step_k(f[0], f[3], DIFFUSION_RATE, f[1], f[2], f[4], f[5], VISCOSITY, DT, kd, kp, True, 1)[:3]
                ok &= judge("published_vs_reference_gpu", glob, ref, 1, dict(n=n, kd=kd, kp=kp, halo=halo, T=T))
            else:
                report.append(dict(case="published_vs_reference_gpu", skipped="libref_gpu.so did not travel to this box"))
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
