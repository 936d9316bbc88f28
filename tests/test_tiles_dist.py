"""The multi-GPU host logic (row-band partition, gather to rank 0, pose sharding) on CPU with gloo, world_size 2."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, band, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    tiles = importlib.import_module("unnamed-voxel-tracer_b200.tiles")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # synthetic "render": pixel value encodes its GLOBAL position, written at the LOCAL storage row
    rpp = tiles.rows_per_part(H, band, world)
    rows = tiles.local_to_global_rows(H, band, world, rank, rpp)
    local = torch.full((rpp, W), -1, dtype=torch.int32)
    for ly, y in enumerate(rows):
        if y >= 0:
            local[ly] = torch.arange(W, dtype=torch.int32) + int(y) * W
    g = tiles.gather_bands(local, 0)
    if rank == 0:
        frame = tiles.assemble(g, H, band)
        np.save(out_path, frame.numpy())
    else:
        assert g is None
    # pose shards partition the sweep exactly
    lo, hi = tiles.shard_poses(256, world, rank)
    t = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(t)
    assert int(t.item()) == 256
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("H,W,band", [(72, 40, 8), (200, 33, 16), (4320 // 8, 16, 32)])
def test_gather_and_assemble_two_ranks(tmp_path, H, W, band):
    import torch.multiprocessing as mp
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, _free_port(), H, W, band, out), nprocs=2, join=True)
    frame = np.load(out)
    assert frame.shape == (H, W)
    assert np.array_equal(frame, np.arange(H * W, dtype=np.int32).reshape(H, W))


@pytest.mark.parametrize("H,band,n", [(1080, 32, 8), (4320, 32, 8), (2160, 16, 8), (4320, 16, 8), (100, 8, 3), (7, 8, 4), (64, 8, 1)])
def test_partition_covers_every_row_exactly_once(uvt, H, band, n):
    t = uvt.tiles
    seen = np.zeros(H, np.int32)
    for part in range(n):
        rows = t.local_to_global_rows(H, band, n, part)
        assert len(rows) == t.storage_rows(H, band, n, part) <= t.rows_per_part(H, band, n)
        seen[rows[rows >= 0]] += 1
    assert (seen == 1).all()
    # interleaving: consecutive bands go to consecutive parts (balances sky and ground rays)
    if n > 1 and H >= band * n:
        owners = [next(p for p in range(n) if b * band in t.local_to_global_rows(H, band, n, p)) for b in range(n)]
        assert owners == list(range(n))


def test_assemble_numpy(uvt):
    t = uvt.tiles
    H, W, band, n = 50, 7, 8, 3
    rpp = t.rows_per_part(H, band, n)
    g = np.full((n, rpp, W), -1, np.int64)
    for p in range(n):
        rows = t.local_to_global_rows(H, band, n, p, rpp)
        for ly, y in enumerate(rows):
            if y >= 0:
                g[p, ly] = y
    f = t.assemble(g, H, band)
    assert np.array_equal(f, np.repeat(np.arange(H)[:, None], W, 1))


def test_shard_poses_balanced(uvt):
    for world in (1, 2, 3, 4, 8):
        spans = [uvt.tiles.shard_poses(256, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 256
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_sweep_poses_are_reproducible_and_inside(uvt):
    a = uvt.scenes.sweep_poses(512, 16)
    b = uvt.scenes.sweep_poses(512, 16)
    assert a.tobytes() == b.tobytes()
    pos = a["cam_pos"][:, :3]
    assert (pos[:, [0, 2]] > 25).all() and (pos[:, [0, 2]] < 487).all() and (pos[:, 1] >= 19).all()
    m = a["cam_mat"].reshape(-1, 4, 4)
    assert np.allclose(np.einsum("nij,nkj->nik", m, m), np.eye(4), atol=1e-5)  # rotations


def test_shared_host_frame_is_one_mapping(uvt):
    """The e2e target of a tiled frame: one POSIX shared-memory frame, every rank writes its own bands."""
    import os
    W, H, band, n = 64, 40, 8, 3
    name = "uvt_cpu_test_%d" % os.getpid()
    a = uvt.tiles.SharedHostFrame(name, W, H, create=True)
    b = uvt.tiles.SharedHostFrame(name, W, H, create=False)
    try:
        a.array[:] = 0
        for part in range(n):
            g = uvt.tiles.local_to_global_rows(H, band, n, part)
            b.array[g[g >= 0]] = part + 1
        want = (np.arange(H) // band) % n + 1
        assert np.array_equal(a.array[:, 0], want) and np.array_equal(a.array[:, -1], want)
    finally:
        b.close()
        a.close()
