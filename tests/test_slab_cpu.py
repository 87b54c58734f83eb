"""Host-side logic of the multi-GPU path on CPU: slab geometry, and a world_size-2 gloo run of the
plumbing bench_multi.py / multi_gpu_check.py use (id broadcast, owned-row reassembly, max-over-ranks
timing).  No compute happens here (there is no CPU fallback to compute with)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world,halo", [(16384, 1, 32), (16384, 2, 32), (16384, 8, 32), (4096, 3, 8), (1000, 7, 16), (32768, 8, 64)])
def test_partition_covers_every_row_exactly_once(f2d, n, world, halo):
    from fluid2d_b200 import slab

    slabs = [slab.partition(n, world, halo, r) for r in range(world)]
    assert slabs[0].own_begin == 0 and slabs[-1].own_end == n
    for a, b in zip(slabs, slabs[1:]):
        assert a.own_end == b.own_begin
    for s in slabs:
        assert s.row_offset == (max(0, s.own_begin - halo) if s.rank > 0 else 0)
        assert s.row_offset + s.rows == (min(n, s.own_end + halo) if s.rank < world - 1 else n)
        b, e = s.local_own
        assert 0 <= b < e <= s.rows and e - b == s.own_end - s.own_begin
        if world > 1:
            assert (b == halo) == (s.rank > 0) and (s.rows - e == halo) == (s.rank < world - 1)
        else:
            assert s.halo == 0 and s.rows == n
    sizes = [s.own_end - s.own_begin for s in slabs]
    assert max(sizes) - min(sizes) <= 1


def test_partition_rejects_slabs_thinner_than_halo(f2d):
    from fluid2d_b200 import slab

    with pytest.raises(ValueError):
        slab.partition(64, 8, 16, 3)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import fluid2d_b200  # noqa: F401
    from fluid2d_b200 import capi, slab

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, halo = 96, 8
    g = np.arange(n * n, dtype=np.float32).reshape(n, n)
    sl = slab.partition(n, world, halo, rank)
    loc = slab.take(sl, g)
    assert loc.shape == (sl.rows, n) and loc[0, 0] == g[sl.row_offset, 0]
    # "compute": each rank stamps its owned rows; halo rows must be ignored by the reassembly
    loc = loc.copy()
    b, e = sl.local_own
    loc[:b] = -1
    loc[e:] = -1
    loc[b:e] += 0.5
    parts = [torch.empty((slab.partition(n, world, halo, r).own_end - slab.partition(n, world, halo, r).own_begin, n)) for r in range(world)]
    dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(loc[b:e])))
    full = torch.cat(parts).numpy()
    assert np.array_equal(full, g + 0.5)
    out = np.zeros_like(g)
    slab.put_owned(sl, out, loc)
    assert np.array_equal(out[sl.own_begin:sl.own_end], g[sl.own_begin:sl.own_end] + 0.5)
    # id broadcast: rank 0's 128 bytes arrive everywhere (the real f2d_comm_unique_id when NCCL loads
    # without a GPU, otherwise a stub with the same signature)
    L = capi.load()
    import ctypes as C
    probe = C.create_string_buffer(128)
    if L.f2d_comm_unique_id(probe) != 0:
        class Stub:
            def __getattr__(self, k):
                return getattr(L, k)

            @staticmethod
            def f2d_comm_unique_id(buf):
                C.memmove(buf, bytes(range(128)), 128)
                return 0
        capi._lib = Stub()
    uid = slab.broadcast_unique_id(dist, rank)
    assert len(uid) == 128
    t = torch.frombuffer(bytearray(uid), dtype=torch.uint8).clone()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref) and int(t.sum()) > 0
    # max-over-ranks timing as in bench_multi.py
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert float(ms) == 10.0 + world - 1
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


def test_world_size_2_gloo_plumbing(f2d, tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(2))
