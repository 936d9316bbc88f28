"""The derivation behind the shadow pass's sun clearance map (csrc/kernels.cuh: top3_kernel / sun_clear_kernel), checked on
the CPU against the oracle's exact traversal: numpy restatement of the two kernels, then tens of thousands of SUN_DIR rays
started ANYWHERE inside empty blocks at rows >= sun1[column] must run into the iteration cap without a hit and without
leaving the map — a far wider set of starts than the surface points the shadow pass shoots from.  (The GPU suite checks
the kernels themselves frame by frame; this file checks the claim they rely on.)"""
import numpy as np
import pytest

WATER = 0x1000000D
SUN = np.float32([7.52185881e-01, 6.58950984e-01, 7.52185881e-01])


def sun_reach_columns(steps):
    return int(np.float32(SUN[0]) * np.float32(steps + 5) / (np.float32(2.0) * SUN[0] + SUN[1])) + 2


def build_maps(tops, steps):
    """top2 = tops grown by the lookup carry ([0, 1]^2); sun1 as sun_clear_kernel computes it."""
    dim = tops.shape[0]                      # tops[z, x]
    pad = np.pad(tops, ((0, 1), (0, 1)), mode="edge")
    top2 = np.maximum.reduce([pad[:-1, :-1], pad[1:, :-1], pad[:-1, 1:], pad[1:, 1:]]).astype(np.int64)
    K = sun_reach_columns(steps)
    climb = np.float32(SUN[1] / SUN[0])
    sun1 = np.zeros((dim, dim), np.int64)
    open_ = np.zeros((dim, dim), bool)
    big = np.pad(top2, ((0, K + 2), (0, K + 2)), constant_values=-1)   # -1 marks columns beyond the +x / +z faces
    for k in range(K + 1):
        credit = max(float(climb * np.float32(max(k - 1, 0)) - np.float32(0.01)), 0.0)
        credit_drift = max(float(climb * np.float32(max(k - 2, 0)) - np.float32(0.01)), 0.0)
        for a, b, cr in ((k, k, credit), (k, k - 1, credit), (k - 1, k, credit), (k + 1, k - 1, credit_drift), (k - 1, k + 1, credit_drift)):
            if a < 0 or b < 0:
                continue
            cell = big[b:b + dim, a:a + dim]
            open_ |= cell < 0
            sun1 = np.maximum(sun1, np.ceil(np.where(cell < 0, 0, cell) - cr).astype(np.int64))
    return top2, np.where(open_, 0xFFFF, np.minimum(sun1, 0xFFFE))


def make_world(uvt, seed, dim=128):
    rng = np.random.default_rng(seed)
    bm = uvt.voxel.VoxelBrickmap.init(dim)
    tops = np.zeros((dim, dim), np.int64)
    for x in range(dim):
        for z in range(dim):
            h = 6 + ((x * (2 + seed % 3) + z) // 5) % 6 + (3 if (x // 7 + z // 9) % 4 == 0 else 0) + int(rng.integers(0, 2))
            for y in range(max(h - 2, 0), h):
                bm.set(x, y, z, int(rng.integers(0, 29)) | (1 << 28))
            tops[z, x] = h
    for _ in range(150):    # floating slabs and pillars
        x, z, y = int(rng.integers(0, dim - 3)), int(rng.integers(0, dim - 3)), int(rng.integers(14, 40))
        for a in range(int(rng.integers(1, 4))):
            for b in range(int(rng.integers(1, 4))):
                for c in range(int(rng.integers(1, 3))):
                    bm.set(x + a, y + c, z + b, WATER)
                    tops[z + b, x + a] = max(tops[z + b, x + a], y + c + 1)
    return bm, tops


@pytest.mark.parametrize("seed,steps", [(0, 48), (1, 48), (2, 17), (3, 96)])
def test_rows_at_or_above_sun1_are_clear_for_the_step_cap(uvt, oracle, atlas, seed, steps):
    dim = 128
    bm, tops = make_world(uvt, seed, dim)
    world = oracle.World(dim, bm.chunks().copy(), bm.bricks().copy(), atlas)
    top2, sun1 = build_maps(tops, steps)
    row_max = dim - 3 - int(np.ceil((steps + 5) * float(SUN[1]) / (2.0 * float(SUN[0]) + float(SUN[1]))))   # uvt.cu: sun_row_max
    rng = np.random.default_rng(100 + seed)
    closed = np.argwhere(sun1 < 0xFFFF)
    assert len(closed) > dim * dim // 3
    n_checked = n_tight = 0
    for i in range(12000):
        z, x = closed[rng.integers(len(closed))]
        s = int(sun1[z, x])
        r = s + int(rng.integers(0, 3))
        if r > row_max:
            continue
        # anywhere in the block, including exactly on its low faces and a hair under its high faces
        f = rng.choice([0.0, 1e-6, 0.125, 0.5, 0.999, 0.999999], 3) if i % 3 == 0 else rng.random(3)
        o = np.float32([x + f[0], r + f[1], z + f[2]])
        if int(o[0]) != x or int(o[1]) != r or int(o[2]) != z:
            continue                       # fp32 rounding pushed the point into the next block
        h = oracle.trace_map(world, o, SUN, steps)
        assert h["data"] == 0 and h["exit_kind"] == 1 and h["trips"] == steps, (seed, x, r, z, o, h)
        n_checked += 1
        if s >= 1 and i % 4 == 0:          # the bound is not vacuous: one row lower, some rays do hit
            o2 = np.float32([x + f[0], s - 1 + f[1], z + f[2]])
            if oracle.trace_map(world, o2, SUN, steps)["data"] != 0:
                n_tight += 1
    assert n_checked > 8000 and n_tight > 100, (n_checked, n_tight)
