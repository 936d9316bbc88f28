"""Host-side formats either side of the traversal path: Voxel word, LCG, VoxelBrickmap +
GpuBlockAllocator, .vox reader, VoxelModelAtlas, procgen, camera."""
import json
import os
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import GOLDEN, REFERENCE_ASSETS


# ---- Voxel word / LCG ---------------------------------------------------------------------
def test_voxel_word_layout(uvt):
    V = uvt.voxel.Voxel
    assert V(13, True) == 0x1000000D           # water, SURVEY A.7(ii)
    assert V(0, True) == 0x10000000            # a valid non-empty block using model 0
    assert V(7) == 7 and V.EMPTY == 0
    assert V(0x0FFFFFFF, True) == 0x1FFFFFFF and V(0x1FFFFFFF) == 0x0FFFFFFF


def test_lcg_stream():
    """util.zig:33-45: seed = seed*1103515245 + 12345 mod 2^32, returns the whole state.
    The native stream is pinned end to end by test_procgen_matches_python_restatement."""
    s, out = 0x46AE4F, []
    for _ in range(3):
        s = (s * 1103515245 + 12345) & 0xFFFFFFFF
        out.append(s)
    assert out == [2270067164, 2712869605, 1569876410]


# ---- VoxelBrickmap ------------------------------------------------------------------------
def test_brickmap_index_formulas(uvt):
    """App. B.1: chunk index cx + cd*(cy + cz*cd); in-brick index lx + 8*ly + 64*lz; entry = brick+1."""
    bm = uvt.voxel.VoxelBrickmap.init(64)
    bm.set(9, 18, 35, 0xABC)
    bm.set(1, 2, 3, 0x123)
    ch, br = bm.chunks(), bm.bricks()
    cd = 8
    assert ch[1 + cd * (2 + 4 * cd)] == 1      # first-touch order: brick 0
    assert ch[0] == 2
    assert br[0][(9 % 8) + 8 * (18 % 8) + 64 * (35 % 8)] == 0xABC
    assert br[1][1 + 8 * 2 + 64 * 3] == 0x123
    assert np.count_nonzero(ch) == 2 and bm.n_bricks == 2
    assert bm.get(9, 18, 35) == 0xABC and bm.get(9, 18, 36) == 0 and bm.get(40, 40, 40) == 0


@settings(max_examples=30, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 31), st.integers(0, 31), st.integers(0, 31), st.integers(1, 2 ** 29 - 1)), min_size=1, max_size=60))
def test_brickmap_set_get_roundtrip(uvt, writes):
    bm = uvt.voxel.VoxelBrickmap.init(32)
    ref = {}
    for x, y, z, v in writes:
        bm.set(x, y, z, v)
        ref[(x, y, z)] = v
    for (x, y, z), v in ref.items():
        assert bm.get(x, y, z) == v
    assert bm.n_bricks == len({(x // 8, y // 8, z // 8) for x, y, z in ref})
    bm.deinit()


def test_allocator_growth_preserves_contents(uvt):
    """gpu_block_allocator.zig:20-30: capacity starts at dim bricks and doubles."""
    bm = uvt.voxel.VoxelBrickmap.init(16)
    assert bm.capacity == 16
    n = 0
    for cx in range(2):
        for cy in range(2):
            for cz in range(2):
                bm.set(cx * 8, cy * 8, cz * 8, 100 + n)
                n += 1
    assert bm.n_bricks == 8 and bm.capacity == 16
    big = uvt.voxel.VoxelBrickmap.init(8 * 5)
    assert big.capacity == 40
    k = 0
    for cx in range(5):
        for cy in range(5):
            for cz in range(5):
                big.set(cx * 8 + 1, cy * 8 + 2, cz * 8 + 3, 1000 + k)
                k += 1
    assert big.n_bricks == 125 and big.capacity == 160  # 40 -> 80 -> 160
    k = 0
    for cx in range(5):
        for cy in range(5):
            for cz in range(5):
                assert big.get(cx * 8 + 1, cy * 8 + 2, cz * 8 + 3) == 1000 + k
                k += 1
    # fresh bricks read as zero
    assert int(big.bricks().sum()) == sum(1000 + i for i in range(125))


def test_is_walkable_and_clear(uvt):
    bm = uvt.voxel.VoxelBrickmap.init(16)
    bm.set(1, 1, 1, uvt.voxel.Voxel(5, True))
    bm.set(2, 1, 1, uvt.voxel.Voxel(7, False))
    assert not bm.is_walkable(1, 1, 1) and bm.is_walkable(2, 1, 1) and bm.is_walkable(3, 1, 1)
    bm.clear()
    assert bm.n_bricks == 0 and bm.get(1, 1, 1) == 0 and not bm.chunks().any()


def test_out_of_range_set_is_an_error_not_a_crash(uvt):
    bm = uvt.voxel.VoxelBrickmap.init(16)
    with pytest.raises(uvt.UvtError):
        bm.set(16, 0, 0, 1)
    with pytest.raises(uvt.UvtError):
        uvt.voxel.VoxelBrickmap.init(12)  # not a multiple of 8


def test_world_dump_roundtrip(uvt, tmp_path, world64):
    path = str(tmp_path / "w.uvtw")
    world64.bm.save(path)
    back = uvt.voxel.VoxelBrickmap.load(path)
    assert back.dim == 64 and back.n_bricks == world64.n_bricks
    assert np.array_equal(back.chunks(), world64.chunks)
    assert np.array_equal(back.bricks(), world64.bricks)


# ---- .vox reader --------------------------------------------------------------------------
def _chunk(cid, content=b"", children=b""):
    return cid + struct.pack("<II", len(content), len(children)) + content + children


def make_vox(models, palette, extra_before_rgba=(), imap=None, version=200):
    """Synthesize a MagicaVoxel file with the chunk structure the assets have."""
    body = b""
    for size, voxels in models:
        body += _chunk(b"SIZE", struct.pack("<III", *size))
        body += _chunk(b"XYZI", struct.pack("<I", len(voxels)) + b"".join(struct.pack("<BBBB", *v) for v in voxels))
    for cid in extra_before_rgba:
        body += _chunk(cid, os.urandom(23))
    body += _chunk(b"RGBA", b"".join(struct.pack("<I", c) for c in palette))
    if imap is not None:
        body += _chunk(b"IMAP", bytes(imap))
    body += _chunk(b"MATL", b"\0" * 40) + _chunk(b"NOTE", b"xyz")
    return b"VOX " + struct.pack("<I", version) + _chunk(b"MAIN", b"", body)


def test_vox_parse_synthetic(uvt):
    pal = [(0xFF000000 | (i * 0x010203)) & 0xFFFFFFFF for i in range(256)]
    m0 = ((8, 8, 8), [(0, 1, 2, 5), (7, 7, 7, 255), (3, 0, 4, 1)])
    m1 = ((8, 8, 8), [(1, 1, 1, 9)])
    data = make_vox([m0, m1], pal, extra_before_rgba=[b"nTRN", b"nGRP", b"nSHP", b"LAYR"], imap=list(range(255, -1, -1)))
    parsed, palette = uvt.voxel.parse_vox(data)
    assert len(parsed) == 2 and parsed[0][0] == (8, 8, 8)
    assert parsed[0][1].tolist() == [list(v) for v in m0[1]]
    assert palette.tolist() == pal  # raw RGBA chunk; IMAP NOT applied (voxel.zig:107 indexes colors[color-1])


def test_vox_parse_rejects_garbage(uvt):
    for bad in (b"", b"VOX ", b"NOPE" + b"\0" * 40, b"VOX " + struct.pack("<I", 150) + b"JUNK" + b"\0" * 20):
        with pytest.raises(uvt.UvtError):
            uvt.voxel.parse_vox(bad)
    # truncated XYZI
    data = make_vox([((8, 8, 8), [(0, 0, 0, 1)] * 4)], [0xFF] * 256)
    with pytest.raises(uvt.UvtError):
        uvt.voxel.parse_vox(data[:60])
    # no palette
    body = _chunk(b"SIZE", struct.pack("<III", 8, 8, 8)) + _chunk(b"XYZI", struct.pack("<I", 0))
    with pytest.raises(uvt.UvtError):
        uvt.voxel.parse_vox(b"VOX " + struct.pack("<I", 200) + _chunk(b"MAIN", b"", body))


def test_atlas_yz_swap_and_slot_order(uvt):
    """voxel.zig:106-108: storage[x + 8*(vox.z + 8*vox.y)] = palette[color-1]; models take consecutive slots."""
    pal = [0xFF000000 + i for i in range(256)]
    m0 = ((8, 8, 8), [(1, 2, 3, 10), (7, 0, 6, 256 - 1)])
    m1 = ((8, 8, 8), [(0, 0, 0, 1)])
    atlas = uvt.voxel.VoxelModelAtlas.init(None)
    atlas.load_block_model(make_vox([m0, m1], pal))
    assert atlas.current_index == 2
    mods = atlas.models()
    exp0 = np.zeros(512, np.uint32)
    exp0[1 + 8 * (3 + 8 * 2)] = pal[9]
    exp0[7 + 8 * (6 + 8 * 0)] = pal[254]
    assert np.array_equal(mods[0], exp0)
    assert mods[1][0] == pal[0] and np.count_nonzero(mods[1]) == 1
    # a model that is not 8^3 is rejected ("assumed to be 8x8x8", voxel.zig:114)
    with pytest.raises(uvt.UvtError):
        atlas.load_block_model(make_vox([((32, 32, 32), [(9, 0, 0, 1)])], pal))


def test_golden_atlas_matches_survey_table(models):
    """filled sub-voxels per model, parsed independently by the surveyor."""
    expect = [385, 416, 447, 457, 446, 443, 49, 78, 53, 40, 34, 176, 46, 384, 342, 229, 234, 320, 331, 331, 331,
              443, 446, 457, 447, 443, 446, 457, 447]
    assert [int((m != 0).sum()) for m in models] == expect
    assert json.load(open(os.path.join(GOLDEN, "atlas_counts.json")))["filled"] == expect
    # all used palette entries are opaque (sub != 0 <=> voxel present)
    assert ((models >> 24)[models != 0] == 0xFF).all()
    # water: top two layers (y' = 6,7) are empty
    water = models[13].reshape(8, 8, 8)  # [z][y][x]
    assert not water[:, 6:, :].any() and (water[:, :6, :] != 0).all()


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ASSETS), reason="reference assets only exist in the build container")
def test_golden_atlas_regenerates_from_reference_assets(uvt, models):
    atlas = uvt.voxel.VoxelModelAtlas.init(None)
    for f in uvt.game.BLOCK_MODEL_FILES:
        atlas.load_block_model(os.path.join(REFERENCE_ASSETS, f))
    assert np.array_equal(atlas.models(), models)


# ---- procgen ------------------------------------------------------------------------------
def _procgen_py(uvt, dim):
    """Pure-Python restatement of procgen.zig:6-70 on a dict world (heights from the native noise)."""
    world = {}
    seed = [0x46AE4F]

    def rand():
        seed[0] = (seed[0] * 1103515245 + 12345) & 0xFFFFFFFF
        return seed[0]

    V = uvt.voxel.Voxel
    for x in range(dim):
        for z in range(dim):
            for y in range(16):
                world[(x, y, z)] = V(13, True)
    for x in range(dim):
        for z in range(dim):
            vh = uvt.procgen.height(dim, x, z)
            for h in range(vh):
                world[(x, h, z)] = V(21 + rand() % 3, True)
                if h <= 15:
                    world[(x, h, z)] = V(25 + rand() % 3, True)
                elif h == vh - 1 and h > 15:
                    world[(x, h, z)] = V(rand() % 6, True)
            if vh > 16:
                if world.get((x, vh, z), 0) != 0:
                    continue
                if rand() % 5 == 0:
                    world[(x, vh, z)] = V(7 + rand() % 5)
                if rand() % 71 == 0:
                    world[(x, vh, z)] = V(12, True)
                if rand() % 420 == 0 and x < 500 and z < 500 and x > 5 and z > 5:
                    th = rand() % 4 + 4
                    world[(x + 1, vh, z + 1)] = V(15, True)
                    for off in range(th):
                        world[(x + 1, vh + off, z + 1)] = V(14 + rand() % 3, True)
                    for a in range(3):
                        for b in range(3):
                            for c in range(3):
                                world[(x + a, vh + th + b, z + c)] = V(18 + rand() % 2, True)
            rand()
    return world


def test_procgen_matches_python_restatement(uvt):
    dim = 512  # full default world: exercises trees, flowers and the x,z<500 guard
    bm = uvt.voxel.VoxelBrickmap.init(dim)
    uvt.procgen.procgen(dim, bm)
    ref = _procgen_py(uvt, dim)
    # compare through the dense block grid
    chunks, bricks = bm.chunks().reshape(64, 64, 64), bm.bricks()  # chunks [cz][cy][cx]
    dense = np.zeros((dim, dim, dim), np.uint32)  # [z][y][x]
    cz, cy, cx = np.nonzero(chunks)
    for z, y, x in zip(cz, cy, cx):
        dense[z * 8:z * 8 + 8, y * 8:y * 8 + 8, x * 8:x * 8 + 8] = bricks[chunks[z, y, x] - 1].reshape(8, 8, 8)
    exp = np.zeros_like(dense)
    for (x, y, z), v in ref.items():
        exp[z, y, x] = v
    assert np.array_equal(dense, exp)
    types = set(np.unique(dense & 0x0FFFFFFF).tolist())
    assert {13, 21, 22, 23, 25, 26, 27, 12, 14, 18}.issubset(types)  # water, dirt, sand, flower, trunk, leaves all occur
    assert ((dense != 0) & ((dense & 0x10000000) == 0)).any()         # non-solid decorations exist


def test_procgen_brick_numbering_is_first_touch_order(uvt, world64):
    """App. B.1: the water slab pass touches chunks x-major, then z, then y=0,1 -> deterministic numbering."""
    ch = world64.chunks.reshape(8, 8, 8)  # [cz][cy][cx]
    assert ch[0, 0, 0] == 1 and ch[0, 1, 0] == 2 and ch[1, 0, 0] == 3 and ch[1, 1, 0] == 4
    assert ch[0, 0, 1] == 17  # next x after 8 z * 2 y


def test_noise_is_bounded_and_smooth(uvt):
    vals = np.array([[uvt.procgen.noise2(x / 10.0, z / 10.0) for z in range(0, 512, 16)] for x in range(0, 512, 16)])
    assert np.abs(vals).max() <= 1.0 and vals.std() > 0.05
    assert np.abs(np.diff(vals, axis=0)).max() < 0.25  # one block = 0.001 simplex units: very smooth
    assert uvt.procgen.noise2(0.0, 0.0) == 0.0         # simplex noise vanishes at lattice origin


# ---- camera -------------------------------------------------------------------------------
def test_camera_uniform_data_layout(uvt):
    cam = uvt.gfx.Camera()
    assert np.float32(cam.fov) == np.float32(np.pi / 2)
    cam.set_pos([1, 2, 3, 0])
    u = cam.as_uniform_data()
    raw = np.frombuffer(u.tobytes(), np.float32)
    assert raw[:4].tolist() == [1, 2, 3, 0]
    assert raw[4:20].reshape(4, 4).tolist() == np.eye(4).tolist() and raw[20] == np.float32(np.pi / 2)
    cam.rotate(100.0, -250.0)  # mouse deltas * 0.001
    assert np.isclose(cam.pitch, 0.1) and np.isclose(cam.yaw, -0.25)
    m = cam.camera_mat()
    fwd = np.array([0, 0, 1, 0], np.float32) @ m  # row vector * matrix
    cp, sp, cy, sy = np.cos(0.1), np.sin(0.1), np.cos(-0.25), np.sin(-0.25)
    assert np.allclose(fwd[:3], [cp * sy, -sp, cp * cy], atol=1e-6)  # positive pitch looks down (App. E.2)
    cam.rotate(1e6, 0)
    assert np.isclose(cam.pitch, np.pi / 2)
    cam.incrementFov(100)
    assert np.isclose(cam.fov, 2.4)


def test_vox_parser_survives_mutations(uvt):
    """Truncations, byte flips and absurd counts in an otherwise valid file end in a parsed file or a UvtError, never
    in a crash or an out-of-bounds read (the parser works on caller memory: uvt_vox_parse(data, size))."""
    rng = np.random.default_rng(7)
    pal = [(0xFF000000 | (i * 0x010203)) & 0xFFFFFFFF for i in range(256)]
    base = make_vox([((8, 8, 8), [(int(a), int(b), int(c), 1 + int(d)) for a, b, c, d in rng.integers(0, 8, (40, 4))]),
                     ((8, 8, 8), [(1, 2, 3, 4)])], pal, extra_before_rgba=[b"nTRN", b"LAYR"], imap=list(range(256)))
    n_ok = n_err = 0
    cases = [base[:k] for k in range(0, len(base), 7)]
    for _ in range(300):
        b = bytearray(base)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        cases.append(bytes(b))
    # chunk sizes and voxel counts far beyond the buffer
    for off in (12, 16, 24, 28, 36, 40, 44):
        b = bytearray(base)
        b[off:off + 4] = struct.pack("<I", 0xFFFFFFF0)
        cases.append(bytes(b))
    for data in cases:
        try:
            models, palette = uvt.voxel.parse_vox(data)
            assert len(palette) == 256
            for size, vox in models:
                assert vox.shape[1] == 4
            n_ok += 1
        except uvt.UvtError:
            n_err += 1
    assert n_ok > 0 and n_err > 0


def test_world_dump_rejects_malformed_files(uvt, tmp_path, world64):
    good = tmp_path / "w.uvtw"
    bm = uvt.voxel.VoxelBrickmap.init(64, 8, None)
    uvt.procgen.procgen(64, bm)
    bm.save(str(good))
    raw = good.read_bytes()
    for name, data in (("magic", b"XXXX" + raw[4:]), ("version", raw[:4] + struct.pack("<I", 9) + raw[8:]),
                       ("truncated_header", raw[:10]), ("truncated_chunks", raw[:16 + 100]), ("truncated_bricks", raw[:-1000]),
                       ("huge_dim", raw[:8] + struct.pack("<I", 0xFFFFFFF8) + raw[12:]), ("odd_dim", raw[:8] + struct.pack("<I", 65) + raw[12:]),
                       ("huge_n_bricks", raw[:12] + struct.pack("<I", 0xFFFFFFFF) + raw[16:]), ("empty", b""),
                       # chunk entries index the brick pool unchecked in get/set: one past the pool, far past it, and a shared brick
                       ("bad_chunk_entry", raw[:16] + struct.pack("<I", bm.n_bricks + 1) + raw[20:]),
                       ("wild_chunk_entry", raw[:16] + struct.pack("<I", 0x20000001) + raw[20:]),
                       ("shared_brick", raw[:16] + raw[16:20] + raw[16:20] + raw[24:])):
        p = tmp_path / f"{name}.uvtw"
        p.write_bytes(data)
        with pytest.raises(uvt.UvtError):
            uvt.voxel.VoxelBrickmap.load(str(p), None)
    with pytest.raises(uvt.UvtError):
        uvt.voxel.VoxelBrickmap.load(str(tmp_path / "missing.uvtw"), None)
    back = uvt.voxel.VoxelBrickmap.load(str(good), None)
    assert back.n_bricks == bm.n_bricks and np.array_equal(back.chunks(), bm.chunks())
