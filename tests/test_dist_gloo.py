"""world_size-2 (and 4) gloo test of the multi-GPU host logic: class partition, in-place all-gather of
the class-planar sub-images and de-interleave reproduce the single-rank image bit for bit. The per-rank
compute is the CPU oracle restricted to the rank's classes (the CUDA equivalent is covered on the GPU by
test_planar_output_and_class_sharding_is_bit_identical)."""
import os
import socket

import numpy as np
import pytest

from mlgi import dist as mdist


def test_class_partition():
    assert mdist.class_range(0, 1) == (0, 16)
    assert [mdist.class_range(r, 2) for r in range(2)] == [(0, 8), (8, 16)]
    assert [mdist.class_range(r, 8) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]
    with pytest.raises(ValueError):
        mdist.class_range(0, 3)
    # class k <-> VPL slots [k*L, (k+1)*L) <-> contiguous ISM rows when L >= S
    for N, S in ((1024, 32), (4096, 64), (16384, 128)):
        L = N // 16
        for world in (2, 4, 8):
            for r in range(world):
                lo, hi = mdist.class_range(r, world)
                assert (lo * L) % S == 0 and (hi * L) % S == 0


def test_row_ownership_and_phases():
    """tile row t belongs to rank t mod world: the ranks' rows partition the image, and every phase of the host-input frame is
    a block of whole tile rows in which each rank owns the same number of rows (+-1 in the last phase)"""
    for H in (216, 2160, 4320, 64, 100):
        TY = (H + 63) // 64
        for world in (2, 4, 8):
            masks = np.stack([mdist.row_mask(r, world, H) for r in range(world)])
            assert (masks.sum(0) == 1).all()
            assert mdist.row_mask(0, world, H)[:min(H, 64)].all()
            for K in (1, 2, 4, 6):
                ph = mdist.host_frame_phases(TY, world, K)
                assert ph[0][0] == 0 and ph[-1][1] == TY and all(a[1] == b[0] for a, b in zip(ph, ph[1:]))
                assert len(ph) <= K
                for (a, b) in ph[:-1]:
                    assert (b - a) % world == 0 and a % world == 0
                for (a, b) in ph:
                    own = [sum(1 for t in range(a, b) if t % world == r) for r in range(world)]
                    assert max(own) - min(own) <= 1


def test_planar_layout_roundtrip():
    rng = np.random.default_rng(0)
    for (H, W) in ((8, 8), (100, 160), (37, 53)):
        img = rng.random((H, W, 4)).astype(np.float32)
        pl = mdist.raster_to_planar(img)
        assert pl.shape == mdist.planar_shape(W, H)
        assert np.array_equal(mdist.planar_to_raster(pl, W, H), img)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, N, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, os.path.join(root, "tests"))
        from conftest import Case
        case = Case(W, H, N, 128, 128, cell=1.5, clutter=6)
        lo, hi = mdist.class_range(rank, world)
        part = case.gather(1, True, bounds=False, class_lo=lo, class_hi=hi)["rgb"]       # this rank's classes only
        rgba = np.concatenate([part, np.zeros((H, W, 1), np.float32)], axis=2)
        planar = torch.zeros(mdist.planar_shape(W, H), dtype=torch.float32)
        planar[lo:hi] = torch.from_numpy(mdist.raster_to_planar(rgba)[lo:hi])
        mdist.allgather_planar(planar, rank, world)
        img = mdist.planar_to_raster(planar.numpy(), W, H)
        if rank == 0:
            full = case.gather(1, True, bounds=False)["rgb"]
            q.put(bool(np.array_equal(img[:, :, :3], full)) and float(np.abs(full).sum()) > 0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_frame_equals_single_rank_bitwise(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 96, 68, 64, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=10) is True
