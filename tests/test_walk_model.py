"""The derivation behind the column-tops walk of the primary pass (csrc/trace.cuh: tops_walk / line_free_trips), checked on
the CPU against the oracle's exact traversal.  A Python restatement of the walk (same pyramid, same margins, same trip bound)
computes n = "the next n trips look up nothing but empty in-map blocks" for tens of thousands of random rays over random
terrain with floating slabs; the oracle then runs those n + 1 trips from the same state and must find no non-empty block,
no map exit — and n >= remaining trips must mean an iteration-cap miss.  The margins of the derivation (4.5 trips, 1/16
block + 0.001 per trip under the line, one block of sideways growth) are what is being tested; the kernel's own fp32
evaluation of the same formulas is covered by the GPU suite (adversarial ray-hugging worlds, full-size frames)."""
import math

import numpy as np
import pytest

from test_sun_clearance_model import make_world


def pyramid(tops):
    dim = tops.shape[0]
    q = dim // 4
    clear4 = np.zeros((q, q), np.int64)
    for qz in range(q):
        for qx in range(q):
            clear4[qz, qx] = tops[max(4 * qz - 1, 0):min(4 * qz + 4, dim - 1) + 1, max(4 * qx - 1, 0):min(4 * qx + 4, dim - 1) + 1].max()

    def coarse(a):
        n = (a.shape[0] + 3) // 4
        out = np.zeros((n, n), np.int64)
        for z in range(n):
            for x in range(n):
                out[z, x] = a[4 * z:4 * z + 4, 4 * x:4 * x + 4].max()
        return out
    clear16 = coarse(clear4)
    return clear4, clear16, coarse(clear16)


def tops_walk(cell, tops, y_all, margin, t0, t_end, p, d, invx, invz, max_iter):
    qdim = tops.shape[0]
    x0, z0 = d[0] * t0 + p[0], d[2] * t0 + p[2]
    qx, qz = int(x0 / cell), int(z0 / cell)          # truncation like the (int) casts of the kernel
    up = d[1] > 0
    sx, sz = (1 if d[0] > 0 else -1), (1 if d[2] > 0 else -1)
    tmx = ((qx + (1 if d[0] > 0 else 0)) * cell - p[0]) * invx
    tmz = ((qz + (1 if d[2] > 0 else 0)) * cell - p[2]) * invz
    tdx, tdz = cell * abs(invx), cell * abs(invz)
    t = t0
    for _ in range(max_iter):
        if not (0 <= qx < qdim and 0 <= qz < qdim):
            break
        t_out = min(tmx, tmz, t_end)
        y_lo = d[1] * (t if up else t_out) + p[1] - margin
        if y_lo < tops[qz, qx]:
            break
        if up and y_lo >= y_all:
            return t_end
        t = t_out
        if t >= t_end:
            break
        if tmx < tmz:
            tmx += tdx
            qx += sx
        else:
            tmz += tdz
            qz += sz
    return t


def line_free_trips(pyr, dim, y_clear, p, d, n_rem):
    clear4, clear16, clear64 = pyr
    invx, invy, invz = 1.0 / d[0], 1.0 / d[1], 1.0 / d[2]
    l1 = abs(d[0]) + abs(d[1]) + abs(d[2])
    t_want = (n_rem + 5) / l1
    hi = dim - 1.0
    ax = ((hi - p[0]) if d[0] > 0 else (p[0] - 1.0)) * abs(invx)
    ay = ((hi - p[1]) if d[1] > 0 else (p[1] - 1.0)) * abs(invy)
    az = ((hi - p[2]) if d[2] > 0 else (p[2] - 1.0)) * abs(invz)
    t_end = min(t_want, ax, ay, az)
    if not t_end > 0:
        return 0
    margin = 0.0625 + 0.001 * (n_rem + 1)
    t_safe = tops_walk(64, clear64, y_clear, margin, 0.0, t_end, p, d, invx, invz, 12)
    if t_safe < t_end:
        t_safe = tops_walk(16, clear16, y_clear, margin, t_safe, t_end, p, d, invx, invz, 32)
    if t_safe < t_end:
        t_safe = tops_walk(4, clear4, y_clear, margin, t_safe, t_end, p, d, invx, invz, 64)
    return min(max(int(t_safe * l1 - 4.5), 0), n_rem)


@pytest.mark.parametrize("seed", range(3))
def test_free_trips_granted_by_the_walk_are_free(uvt, oracle, atlas, seed):
    dim = 128
    bm, tops = make_world(uvt, 10 + seed, dim)
    world = oracle.World(dim, bm.chunks().copy(), bm.bricks().copy(), atlas)
    pyr = pyramid(tops)
    y_clear = int(pyr[0].max())
    rng = np.random.default_rng(500 + seed)
    granted = sealed = total = 0
    for i in range(9000):
        x, z = rng.uniform(1.5, dim - 1.5, 2)
        y = tops[int(z), int(x)] + rng.choice([0.05, 0.3, 1.0, 2.5, 6.0, 20.0, 60.0]) * rng.random()
        if y >= dim - 1:
            continue
        d = rng.normal(size=3)
        if i % 3 == 0:
            d[1] = abs(d[1]) * rng.choice([0.02, 0.2, 1.0])      # climbing, some of them grazing
        elif i % 3 == 1:
            d[1] = -abs(d[1]) * rng.choice([0.02, 0.1, 0.5])     # descending at shallow angles over the terrain
        d = (d / np.linalg.norm(d)).astype(np.float32)
        if (np.abs(d) < 1e-4).any():
            continue
        o = np.float32([x, y, z])
        if world_block_nonempty(bm, o):
            continue
        trip0 = int(rng.integers(0, 180))
        n_rem = 192 - trip0 - 1
        n = line_free_trips(pyr, dim, y_clear, [float(v) for v in o], [float(v) for v in d], n_rem)
        total += 1
        if n == 0:
            continue
        granted += 1
        h = oracle.trace_map(world, o, d, n + 1)          # the current trip's lookup + the n that follow
        assert h["t_block"] == 0 and h["data"] == 0 and h["exit_kind"] == 1 and h["trips"] == n + 1, (seed, o, d, n, n_rem, h)
        if n >= n_rem:
            sealed += 1
    assert granted > total // 8 and sealed > 200, (total, granted, sealed)


def world_block_nonempty(bm, o):
    return bm.get(int(o[0]), int(o[1]), int(o[2])) != 0


def test_block_clearance_bound(uvt, oracle, atlas):
    """The free-trip bound of the dense grid: with D the Chebyshev distance (blocks) from the looked-up empty block
    to the nearest non-empty block or map face, the D - 2 trips that follow look up empty in-map blocks, whatever the direction."""
    dim = 64
    rng = np.random.default_rng(77)
    bm = uvt.voxel.VoxelBrickmap.init(dim)
    occ = np.zeros((dim, dim, dim), bool)   # [x, y, z]
    for _ in range(260):
        x, y, z = (int(v) for v in rng.integers(0, dim, 3))
        bm.set(x, y, z, int(rng.integers(0, 29)) | (1 << 28))
        occ[x, y, z] = True
    world = oracle.World(dim, bm.chunks().copy(), bm.bricks().copy(), atlas)
    pts = np.argwhere(occ)
    checked = 0
    for i in range(6000):
        b = rng.integers(0, dim, 3)
        if occ[tuple(b)]:
            continue
        D = int(np.abs(pts - b).max(axis=1).min())
        D = min(D, int(min(b.min() + 1, (dim - b).min())), 16)
        if D < 3:
            continue
        f = rng.choice([0.0, 1e-6, 0.5, 0.999999], 3) if i % 2 else rng.random(3)
        o = (b + f).astype(np.float32)
        if (o.astype(np.int64) != b).any():
            continue
        d = rng.normal(size=3)
        d = (d / np.linalg.norm(d)).astype(np.float32)
        if i % 5 == 0:
            d[rng.integers(3)] = 0.0
        n = D - 2
        h = oracle.trace_map(world, o, d, n + 1)
        assert h["t_block"] == 0 and h["data"] == 0 and h["exit_kind"] == 1, (b, D, o, d, h)
        checked += 1
    assert checked > 2500
