"""Row-slab decomposition for the multi-GPU path (one process per GPU, torch.distributed plumbing).

The global N-row grid is cut into `world` contiguous blocks of rows (global edge rows 0 and N-1
belong to the first / last block).  Rank r holds its block plus `halo` rows of each neighbour; the
CUDA library exchanges those rows over NCCL (f2d_comm_*, include/f2d.h).  Nothing here computes:
this module only derives slab geometry, slices host arrays and wires the communicator.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .solver import FluidSolverB200


@dataclass(frozen=True)
class Slab:
    rank: int
    world: int
    global_rows: int
    halo: int
    own_begin: int   # global row range owned by this rank: [own_begin, own_end)
    own_end: int
    row_offset: int  # global row of local row 0
    rows: int        # local rows including halos

    @property
    def local_own(self):
        """Owned rows in local coordinates."""
        return self.own_begin - self.row_offset, self.own_end - self.row_offset


def partition(global_rows, world, halo, rank):
    """Slab of `rank`.  Blocks differ by at most one row; every block must be >= 3 halos deep."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(global_rows, world)
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    lo = max(0, begin - halo) if rank > 0 else 0
    hi = min(global_rows, end + halo) if rank < world - 1 else global_rows
    if world > 1 and (end - begin) < max(1, halo):
        raise ValueError("slab of %d rows is thinner than the halo (%d)" % (end - begin, halo))
    return Slab(rank, world, global_rows, halo if world > 1 else 0, begin, end, lo, hi - lo)


def take(slab, a):
    """Local slab (halo rows included) of a global host array, as an independent COPY (a contiguous row range of a
    C-ordered array would otherwise be a view, and solve() updates its grids in place)."""
    return np.array(a[slab.row_offset:slab.row_offset + slab.rows], dtype=np.float32, order="C", copy=True)


def put_owned(slab, dst, local):
    """Write the owned rows of a local slab back into a global host array."""
    b, e = slab.local_own
    dst[slab.own_begin:slab.own_end] = local[b:e]


def broadcast_unique_id(dist, rank, device=None):
    """Rank 0 creates the 128-byte NCCL id (f2d_comm_unique_id); torch.distributed broadcasts it."""
    import torch

    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        raw = C.create_string_buffer(128)
        capi.check(capi.load().f2d_comm_unique_id(raw))
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def gather_p2p_handles(dist, solver, world, device=None):
    """Every rank exports the CUDA IPC handle of its arena + 4 words of geometry (f2d_p2p_export);
    torch.distributed all-gathers the 96-byte records.  Returns a list of (handle bytes, info array)."""
    import torch

    h = (C.c_ubyte * 64)()
    info = (C.c_uint64 * 4)()
    rc = capi.load().f2d_p2p_export(solver._h, h, info)  # a failing rank still takes part in the all-gather
    if rc != 0:
        info = (C.c_uint64 * 4)(0, 0, 0, 0)
    rec = torch.frombuffer(bytearray(bytes(h) + bytes(info)), dtype=torch.uint8).clone()
    if device is not None:
        rec = rec.to(device)
    out = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(out, rec)
    recs = []
    for t in out:
        b = t.cpu().numpy().tobytes()
        recs.append((b[:64], np.frombuffer(b[64:96], dtype=np.uint64).copy()))
    if any(int(r[1][3]) == 0 for r in recs):  # same verdict on every rank
        raise capi.F2DError(capi.ERR_CUDA, "a rank could not export its CUDA IPC handle")
    return recs


def make_slab_solver(slab, cols, unique_id=None, cfl_cells=8, device=0, transport="auto", dist=None, torch_device=None,
                     **solver_kwargs):
    """FluidSolverB200 for one slab with its halo transport wired (collective over all ranks).
    transport="p2p" : direct peer stores over NVLink (needs dist: IPC handles are all-gathered here);
    transport="nccl": NCCL send/recv pairs (needs unique_id from broadcast_unique_id);
    transport="auto": p2p when every rank can map its neighbours, else nccl (collective decision)."""
    s = FluidSolverB200(slab.rows, cols, global_rows=slab.global_rows, row_offset=slab.row_offset,
                        halo=slab.halo, device=device, **solver_kwargs)
    if slab.world > 1:
        L = capi.load()
        if transport == "auto":
            # peer-memory transport when CUDA IPC works on every rank, else NCCL; the decision is collective
            import torch

            ok = 1
            try:
                _connect_p2p(L, s, slab, dist, torch_device, cfl_cells)
            except capi.F2DError:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=torch_device if torch_device is not None else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                dist.barrier()
                s.transport = "p2p"
                return s
            s.close()  # some rank could not map its neighbours: everybody rebuilds with NCCL
            s = FluidSolverB200(slab.rows, cols, global_rows=slab.global_rows, row_offset=slab.row_offset,
                                halo=slab.halo, device=device, **solver_kwargs)
            unique_id = broadcast_unique_id(dist, slab.rank, device=torch_device)
            capi.check(L.f2d_comm_init(s._h, unique_id, slab.rank, slab.world, cfl_cells))
            s.transport = "nccl"
            return s
        if transport == "p2p":
            _connect_p2p(L, s, slab, dist, torch_device, cfl_cells)
            dist.barrier()  # every rank has opened its neighbours' arenas before anyone steps
            s.transport = "p2p"
            return s
        capi.check(L.f2d_comm_init(s._h, unique_id, slab.rank, slab.world, cfl_cells))
        s.transport = "nccl"
    return s


def _connect_p2p(L, s, slab, dist, torch_device, cfl_cells):
    recs = gather_p2p_handles(dist, s, slab.world, device=torch_device)

    def arg(r):
        if r < 0 or r >= slab.world:
            return None, None
        hb, info = recs[r]
        return (C.c_ubyte * 64).from_buffer_copy(hb), (C.c_uint64 * 4)(*[int(x) for x in info])

    uh, ui = arg(slab.rank - 1)
    dh, di = arg(slab.rank + 1)
    capi.check(L.f2d_p2p_connect(s._h, slab.rank, slab.world, uh, ui, dh, di, cfl_cells))


def comm_exchanges(solver):
    n = C.c_uint64()
    capi.check(capi.load().f2d_comm_stats(solver._h, C.byref(n)))
    return int(n.value)


def comm_bytes(solver):
    """Payload bytes this rank has pushed to ONE neighbour so far (f2d_comm_bytes)."""
    up, down = C.c_uint64(), C.c_uint64()
    capi.check(capi.load().f2d_comm_bytes(solver._h, C.byref(up), C.byref(down)))
    return max(int(up.value), int(down.value))
