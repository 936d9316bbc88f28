"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): hit voxel coordinates, face ids, material (block word) and colour
BIT-EXACT; G-buffer normal / position / illumination bit-exact; hit distance within 2 ulp; sky
albedo and the shaded frame within 1/255 per channel.
"""
import numpy as np
import pytest

from conftest import camera_k0, camera_k1, pitch_yaw_matrix

pytestmark = pytest.mark.gpu

WATER = 0x1000000D


def channel_diff(a, b):
    a8 = a.view(np.uint8).astype(np.int16)
    b8 = b.view(np.uint8).astype(np.int16)
    return np.abs(a8 - b8)


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


def assert_primary_parity(gpu, ref):
    """gpu/ref: dicts with albedo, normal, position, hits."""
    h, g = ref["hits"], gpu["hits"]
    for f in ("px", "py", "pz", "block", "color", "face", "trips", "exit_kind"):
        assert np.array_equal(g[f], h[f]), f"hit buffer field {f} differs at {np.argwhere(g[f] != h[f])[:5]}"
    hit = h["face"] != 0
    assert ulp_diff(g["distance"][hit], h["distance"][hit]).max(initial=0) <= 2
    assert (g["distance"][~hit] == -1.0).all()
    assert np.array_equal(gpu["normal"], ref["normal"])
    assert np.array_equal(gpu["position"].view(np.uint32), ref["position"].view(np.uint32))
    assert np.array_equal(gpu["albedo"][hit], ref["albedo"][hit])          # colours of hits: bit-exact
    assert channel_diff(gpu["albedo"], ref["albedo"]).max() <= 1            # sky: 1/255


def gpu_render(ctx, cam, three_pass=True):
    ctx.set_camera(cam)
    if three_pass:
        ctx.dispatch_primary()
        ctx.dispatch_secondary()
        ctx.shade()
    else:
        ctx.dispatch_frame()
    out = {k: ctx.readback(k) for k in ("albedo", "normal", "position", "illumination", "frame")}
    try:
        out["hits"] = ctx.readback("hit")
    except Exception:
        out["hits"] = None
    return out


@pytest.fixture(scope="module")
def w1(uvt, scene_factory):
    """W1 on the GPU: procgen(512) written into the ctx's pinned staging and committed through the C ABI."""
    ctx = uvt.Context(0, hit_buffer=True)
    sc = scene_factory(512, "procgen", ctx=ctx)
    yield ctx, sc
    ctx.close()


@pytest.mark.parametrize("layout", ["compact", "reference"])
@pytest.mark.parametrize("cam_name,size", [("k0", (320, 180)), ("k1", (320, 180)), ("k1", (1280, 720))])
def test_frame_parity_w1(uvt, oracle, w1, layout, cam_name, size):
    ctx, sc = w1
    W, H = size
    cam = camera_k0(oracle) if cam_name == "k0" else camera_k1(uvt, oracle)
    ctx.set_layout(layout)
    assert ctx.effective_layout() == layout
    ctx.resize(W, H)
    g = gpu_render(ctx, cam)
    r = oracle.render(sc.oracle_world, cam, W, H)
    assert_primary_parity(g, r)
    assert np.array_equal(g["illumination"], r["illumination"])
    assert channel_diff(g["frame"], r["frame"]).max() <= 1
    assert (r["hits"]["face"] != 0).mean() > 0.3 and len(np.unique(r["illumination"])) == 3
    # exact traversal counters (they define the algorithmic bytes of the roofline)
    assert ctx.count_pass("primary") == {**r["primary_counters"]}
    assert ctx.count_pass("secondary") == {**r["secondary_counters"]}


def test_committed_golden_fixture(uvt, oracle, w1):
    """The GPU against the committed oracle output (tests/golden/oracle_w1_*.npz)."""
    import os
    from conftest import GOLDEN
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(96, 54)
    for name, cam in (("k0_96x54", camera_k0(oracle)), ("k1_96x54", camera_k1(uvt, oracle))):
        gold = np.load(os.path.join(GOLDEN, f"oracle_w1_{name}.npz"))
        g = gpu_render(ctx, cam)
        gh = gold["hits"].view(g["hits"].dtype).reshape(54, 96)
        for f in ("px", "py", "pz", "block", "color", "face", "trips", "exit_kind"):
            assert np.array_equal(g["hits"][f], gh[f]), f
        hit = gh["face"] != 0
        assert ulp_diff(g["hits"]["distance"][hit], gh["distance"][hit]).max(initial=0) <= 2 and (g["hits"]["distance"][~hit] == -1.0).all()
        assert np.array_equal(g["albedo"][hit], gold["albedo"][hit]) and channel_diff(g["albedo"], gold["albedo"]).max() <= 1
        for k in ("normal", "illumination"):
            assert np.array_equal(g[k], gold[k])
        assert np.array_equal(g["position"].view(np.uint32), gold["position"].view(np.uint32))
        assert channel_diff(g["frame"], gold["frame"]).max() <= 1


def test_dispatch_frame_equals_three_passes(uvt, oracle, w1):
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(640, 360)
    cam = camera_k1(uvt, oracle)
    a = gpu_render(ctx, cam, three_pass=True)
    b = gpu_render(ctx, cam, three_pass=False)
    ctx.dispatch_primary()
    ctx.dispatch_secondary_shade()     # the shadow pass and the blit in one launch
    for k in ("illumination", "frame"):
        assert np.array_equal(a[k], ctx.readback(k)), k
    for k in ("albedo", "normal", "illumination", "frame"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["position"].view(np.uint32), b["position"].view(np.uint32))


@pytest.mark.parametrize("variant", ["bricks", "pool", "pool-bricks", "fused", "fused-bricks"])
def test_every_kernel_variant_matches_the_oracle(uvt, oracle, scene_factory, variant):
    """The traversal exists in several instantiations (dense block grid / chunk table + bricks; pixel-per-thread /
    pooled scheduler; three launches / fused frame kernel).  The default (dense, tile, three launches) is what every
    other test runs; here each alternative is held to the same bit-exact bar."""
    kw = dict(hit_buffer=True, dense="bricks" not in variant, fused_frame=variant.startswith("fused"))
    with uvt.Context(0, **kw) as ctx:
        sc = scene_factory(512, "procgen", ctx=ctx)
        ctx.set_scheduler("pool" if variant.startswith("pool") else "tile")
        for cam, size in ((camera_k1(uvt, oracle), (320, 180)), (camera_k0(oracle), (250, 130))):
            ctx.resize(*size)
            g = gpu_render(ctx, cam, three_pass=not variant.startswith("fused"))
            r = oracle.render(sc.oracle_world, cam, *size)
            if variant.startswith("fused"):  # the fused kernel does not write the explicit hit buffer: G-buffer parity
                hit = r["hits"]["face"] != 0
                assert np.array_equal(g["normal"], r["normal"])
                assert np.array_equal(g["position"].view(np.uint32), r["position"].view(np.uint32))
                assert np.array_equal(g["albedo"][hit], r["albedo"][hit]) and channel_diff(g["albedo"], r["albedo"]).max() <= 1
            else:
                assert_primary_parity(g, r)
            assert np.array_equal(g["illumination"], r["illumination"])
            assert channel_diff(g["frame"], r["frame"]).max() <= 1
            assert ctx.count_pass("primary") == r["primary_counters"]
            assert ctx.count_pass("secondary") == r["secondary_counters"]
        if not variant.startswith("fused"):  # a few random poses against the verbatim-trips kernel, GPU vs GPU
            rng = np.random.default_rng(7)
            ctx.resize(192, 108)
            for i in range(10):
                x, z = rng.uniform(20, 490, 2)
                cam = oracle.make_camera((x, uvt.procgen.height(512, int(x), int(z)) + rng.uniform(2, 80), z),
                                         pitch_yaw_matrix(uvt, rng.uniform(-1.2, 1.2), rng.uniform(0, 6.28)), fov=rng.uniform(0.5, 2.2))
                ctx.set_layout("compact")
                a = gpu_render(ctx, cam)
                ctx.set_layout("reference")
                b = gpu_render(ctx, cam)
                ctx.set_layout("compact")
                assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8)), (variant, i)
                assert np.array_equal(a["illumination"], b["illumination"]) and np.array_equal(a["frame"], b["frame"])


@pytest.mark.parametrize("size", [(1, 1), (17, 9), (33, 31), (250, 130)])
def test_ragged_sizes(uvt, oracle, w1, size):
    """Sizes that are not multiples of the 16x8 CTA tile (the reference over-dispatches and bounds-checks, primary.comp.glsl:28-29)."""
    ctx, sc = w1
    W, H = size
    ctx.resize(W, H)
    cam = camera_k1(uvt, oracle)
    g = gpu_render(ctx, cam)
    r = oracle.render(sc.oracle_world, cam, W, H)
    assert_primary_parity(g, r)
    assert np.array_equal(g["illumination"], r["illumination"])
    assert channel_diff(g["frame"], r["frame"]).max() <= 1


def test_random_poses_hit_buffers_bit_exact(uvt, oracle, w1):
    """Random interior poses (the C5 pose distribution), hit buffers only."""
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(160, 90)
    rng = np.random.default_rng(5)
    for i in range(12):
        x, z = rng.uniform(20, 490, 2)
        y = uvt.procgen.height(512, int(x), int(z)) + rng.uniform(3, 40)
        cam = oracle.make_camera((x, y, z), pitch_yaw_matrix(uvt, rng.uniform(-0.6, 0.6), rng.uniform(0, 2 * np.pi)),
                                 fov=rng.uniform(0.6, 2.2))
        g = gpu_render(ctx, cam)
        r = oracle.render(sc.oracle_world, cam, 160, 90)
        assert_primary_parity(g, r)
        assert np.array_equal(g["illumination"], r["illumination"])


def test_camera_outside_and_degenerate_views(uvt, oracle, w1):
    """Camera outside the map (AABB clip path), looking straight down/up (zero direction components)."""
    ctx, sc = w1
    ctx.resize(128, 72)
    down = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float32)
    cams = [oracle.make_camera((-40.0, 60.0, 256.0), pitch_yaw_matrix(uvt, 0.3, np.pi / 2)),   # outside, looking in (+x)
            oracle.make_camera((256.0, 700.0, 256.0), down),                                    # far above, looking down
            oracle.make_camera((256.0, 30.0, 256.0), down),
            oracle.make_camera((600.0, 30.0, 600.0), None),                                      # outside, looking away
            oracle.make_camera((256.0, 25.0, 256.0), np.eye(4, dtype=np.float32), fov=0.314)]
    for cam in cams:
        g = gpu_render(ctx, cam)
        r = oracle.render(sc.oracle_world, cam, 128, 72)
        assert_primary_parity(g, r)
        assert np.array_equal(g["illumination"], r["illumination"])


def test_entity_boxes_shadow_as_lines(uvt, oracle, w1):
    """traceEntities (map.glsl:172-201): the five unit boxes shadow as LINES, boxes behind the origin included.  Two cameras whose
    frames hold many pixels that only the entities put in shadow (oracle with entities on vs off), against the GPU."""
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(256, 144)
    down = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float32)
    for cam, least in ((oracle.make_camera((256.0, 30.0, 256.0), down), 150),
                       (oracle.make_camera((262.0, 30.0, 262.0), pitch_yaw_matrix(uvt, 0.6, 5 * np.pi / 4)), 300),
                       (oracle.make_camera((249.0, 27.0, 262.0), pitch_yaw_matrix(uvt, 0.9, 2.2)), 1)):
        g = gpu_render(ctx, cam)
        r = oracle.render(sc.oracle_world, cam, 256, 144)
        off = oracle.render(sc.oracle_world, cam, 256, 144, oracle.params(512, entities=False))
        n_entity = int((r["illumination"] != off["illumination"]).sum())
        assert n_entity >= least, n_entity
        assert np.array_equal(g["illumination"], r["illumination"])
        assert_primary_parity(g, r)
        assert channel_diff(g["frame"], r["frame"]).max() <= 1


def test_empty_world_and_single_block(uvt, oracle, scene_factory):
    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(512, None, ctx=ctx)
        ctx.resize(64, 36)
        g = gpu_render(ctx, camera_k0(oracle))
        r = oracle.render(sc.oracle_world, camera_k0(oracle), 64, 36)
        assert_primary_parity(g, r)
        assert (g["hits"]["face"] == 0).all() and (g["illumination"] == 0).all() and g["hits"]["trips"].max() == 192
    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(512, lambda bm: bm.set(256, 0, 256, WATER), ctx=ctx)
        ctx.resize(64, 64)
        down = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float32)
        cam = oracle.make_camera((256.5, 4.0, 256.5), down)
        g = gpu_render(ctx, cam)
        rec = g["hits"][32, 32]
        assert (rec["px"], rec["py"], rec["pz"], rec["face"], rec["trips"], rec["color"], rec["block"]) == (2052, 5, 2052, 4, 6, 0xFFFFCC99, WATER)
        assert g["position"][32, 32].tolist() == [256.625, 0.75, 256.625, 1.0]
        assert_primary_parity(g, oracle.render(sc.oracle_world, cam, 64, 64))


def test_small_world_dims(uvt, oracle, scene_factory):
    """MAP_DIMENSION other than 512 (the C3 world uses 2048): bounds and chunk-table strides follow dim."""
    for dim in (64, 128):
        with uvt.Context(0, hit_buffer=True, map_dim=dim) as ctx:
            sc = scene_factory(dim, "procgen", ctx=ctx)
            ctx.resize(96, 54)
            cam = oracle.make_camera((dim / 2, 22.0, dim / 2), pitch_yaw_matrix(uvt, 0.4, 0.8))
            g = gpu_render(ctx, cam)
            r = oracle.render(sc.oracle_world, cam, 96, 54, oracle.params(dim))
            assert_primary_parity(g, r)
            assert np.array_equal(g["illumination"], r["illumination"])


def test_unloaded_model_reads_empty_and_many_materials_fall_back(uvt, oracle, models, atlas, scene_factory):
    """A block whose model slot was never uploaded reads as empty; > 255 distinct block words
    cannot use the 8-bit compact bricks and must fall back to the reference layout, still bit-exact."""
    def fill(bm):
        for x in range(240, 272):
            for z in range(240, 272):
                bm.set(x, 0, z, WATER)
                bm.set(x, 1, z, 0x10000000 | 200)          # model 200: not loaded
        for i in range(300):
            # 300 distinct block words (bits 15.. vary) that all resolve to loaded model slots (word & 32767 = i % 29)
            bm.set(240 + i % 32, 3, 240 + i // 32, (i % 29) | 0x10000000 | ((i + 1) << 15))
    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(512, fill, ctx=ctx)
        assert ctx.effective_layout() == "reference"
        ctx.resize(96, 54)
        cam = oracle.make_camera((256.0, 12.0, 250.0), pitch_yaw_matrix(uvt, 0.9, 0.1))
        g = gpu_render(ctx, cam)
        r = oracle.render(sc.oracle_world, cam, 96, 54)
        assert_primary_parity(g, r)
        assert (r["hits"]["face"] != 0).any()


def test_partition_union_equals_full_frame(uvt, oracle, w1):
    """Multi-GPU image partition: interleaved row bands rendered separately reassemble to the full frame."""
    ctx, sc = w1
    ctx.set_layout("compact")
    W, H = 320, 200  # 25 bands of 8 rows: ragged against 4 parts
    cam = camera_k1(uvt, oracle)
    ctx.set_partition(8, 1, 0)
    ctx.resize(W, H)
    full = gpu_render(ctx, cam)
    n_parts, band = 4, 16
    frame = np.zeros((H, W), np.uint32)
    pos = np.zeros((H, W, 4), np.float32)
    for part in range(n_parts):
        ctx.set_partition(band, n_parts, part)
        g = gpu_render(ctx, cam)
        rows = ctx.local_rows()
        assert g["frame"].shape == (rows, W)
        for ly in range(rows):
            lb = ly // band
            y = (lb * n_parts + part) * band + ly % band
            if y < H:
                frame[y] = g["frame"][ly]
                pos[y] = g["position"][ly]
    ctx.set_partition(8, 1, 0)
    assert np.array_equal(frame, full["frame"])
    assert np.array_equal(pos.view(np.uint32), full["position"].view(np.uint32))


def test_band_readback_assembles_the_frame_in_host_memory(uvt, oracle, w1):
    """uvt_readback_bands_async: every part copies its bands straight to their rows of ONE host frame (the e2e path of a
    tiled frame: N parallel device-to-host copies, no GPU-to-GPU hop) — ragged last band included."""
    ctx, sc = w1
    ctx.set_layout("compact")
    W, H = 320, 200  # 13 bands of 16 rows, the last one 8 rows
    cam = camera_k1(uvt, oracle)
    ctx.set_partition(8, 1, 0)
    ctx.resize(W, H)
    full = gpu_render(ctx, cam)["frame"]
    shm = uvt.tiles.SharedHostFrame("uvt_test_frame_%d" % __import__("os").getpid(), W, H, create=True)
    try:
        shm.array[:] = 0
        ctx.host_register(shm.array)
        for n_parts in (1, 3, 4):
            shm.array[:] = 0xDEADBEEF
            for part in range(n_parts):
                ctx.set_partition(16, n_parts, part)
                ctx.set_camera(cam)
                ctx.dispatch_frame()
                ctx.readback_bands_async(shm.array)
            ctx.readback_wait()
            assert np.array_equal(shm.array, full), n_parts
        ctx.host_unregister(shm.array)
    finally:
        ctx.set_partition(8, 1, 0)
        shm.close()


def test_batched_poses_equal_single_dispatches(uvt, oracle, w1):
    ctx, sc = w1
    ctx.resize(160, 90)
    cams = np.stack([camera_k0(oracle), camera_k1(uvt, oracle), oracle.make_camera((100.0, 40.0, 300.0), pitch_yaw_matrix(uvt, 0.2, 2.0))])
    singles = [gpu_render(ctx, c) for c in cams]
    ctx.set_camera(cams)
    ctx.dispatch_primary(); ctx.dispatch_secondary(); ctx.shade()
    fr, hb = ctx.readback("frame"), ctx.readback("hit")
    assert fr.shape == (3, 90, 160)
    for i in range(3):
        assert np.array_equal(fr[i], singles[i]["frame"])
        assert np.array_equal(hb[i].view(np.uint8), singles[i]["hits"].view(np.uint8))
    ctx.set_camera(cams[0])


def test_pick_ray_matches_oracle(uvt, oracle, w1):
    """terrain_edit.comp.glsl: the centre ray with a 64-trip cap."""
    ctx, sc = w1
    cam = camera_k1(uvt, oracle)
    ctx.set_camera(cam)
    rec = ctx.pick()
    # centre ray = uv 0: same as pixel (W/2, H/2) of any even-sized frame
    o, d, s = oracle.primary_ray(cam, 2, 2, 1, 1, 512)
    h = oracle.trace_map(sc.oracle_world, s, d, 64)
    assert (int(rec["px"]), int(rec["py"]), int(rec["pz"])) == h["p"]
    assert (int(rec["face"]), int(rec["block"]), int(rec["color"]), int(rec["trips"])) == (h["face"], h["block"], h["data"], h["trips"])


def test_game_mirror_call_order(uvt, oracle, models, w1):
    """src/game.zig:54-129,224-256 replayed through the gfx mirror: same result as the direct C calls."""
    with uvt.Context(0) as ctx:
        game = uvt.game.Game(ctx, dim=512, width=320, height=180, models=models)
        game.update()
        game.pre_render()
        game.render()
        frame = ctx.readback("frame")
        _, sc = w1
        r = oracle.render(sc.oracle_world, camera_k0(oracle), 320, 180)
        assert channel_diff(frame, r["frame"]).max() <= 1
        assert np.array_equal(ctx.readback("illumination"), r["illumination"])
        with pytest.raises(uvt.UvtError):
            game.primary_trace_pipeline.dispatch(1, 1, 1)  # does not cover the G-buffer
        game.window_resized(100, 60)
        game.render()
        assert ctx.readback("frame").shape == (60, 100)
        game.deinit()


def test_errors_are_reported_not_fatal(uvt, oracle):
    with uvt.Context(0) as ctx:
        with pytest.raises(uvt.UvtError):
            ctx.dispatch_primary()           # no G-buffer
        ctx.resize(32, 32)
        with pytest.raises(uvt.UvtError):
            ctx.dispatch_primary()           # no camera
        ctx.set_camera(camera_k0(oracle))
        with pytest.raises(uvt.UvtError) as e:
            ctx.dispatch_primary()           # no world
        assert "world" in e.value.message
        with pytest.raises(uvt.UvtError):
            ctx.readback("hit")              # hit buffer not enabled
        with pytest.raises(uvt.UvtError):
            ctx.set_partition(5, 2, 0)       # band not a multiple of 8
        with pytest.raises(uvt.UvtError):
            ctx.set_entities(np.zeros((33, 3), np.float32))          # more than 32 entities
        with pytest.raises(uvt.UvtError):
            ctx.entity_model_upload(np.zeros(12 ** 3, np.uint32), 12)  # edge not 8 / 16 / 32
        with pytest.raises(uvt.UvtError):
            ctx.set_frame_chunks(0)
        with pytest.raises(uvt.UvtError):
            ctx.check(ctx.L.uvt_set_entity_mode(ctx.handle, 7))
        ctx.set_entity_mode("models")        # turns the hit buffer on: the G-buffer is re-created with one
        assert ctx.readback("hit").shape == (32, 32)


def test_world_edit_then_recommit(uvt, oracle, scene_factory):
    """VoxelBrickmap.set after init is visible after the next bind() (the reference mapping is live)."""
    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(512, "procgen", ctx=ctx)
        ctx.resize(160, 90)
        cam = camera_k0(oracle)
        before = gpu_render(ctx, cam)
        for y in range(20, 40):
            for x in range(250, 262):
                sc.bm.set(x, y, 270, uvt.voxel.Voxel(11, True))   # a rock wall in front of the camera, allocates new bricks
        sc.bm.bind(9)
        after = gpu_render(ctx, cam)
        world = oracle.World(512, sc.bm.chunks().copy(), sc.bm.bricks().copy(), sc.oracle_world.atlas)
        r = oracle.render(world, cam, 160, 90)
        assert_primary_parity(after, r)
        assert not np.array_equal(before["hits"]["block"], after["hits"]["block"])


# ---- full-size, size-independent properties (BASELINE configs) -------------------------------------
@pytest.mark.parametrize("size", [(1920, 1080), (3840, 2160)])
def test_full_size_layouts_agree_and_rows_match_oracle(uvt, oracle, w1, size):
    ctx, sc = w1
    W, H = size
    cam = camera_k1(uvt, oracle)
    ctx.resize(W, H)
    ctx.set_layout("compact")
    a = gpu_render(ctx, cam)
    ca = ctx.count_pass("primary")
    ctx.set_layout("reference")
    b = gpu_render(ctx, cam)
    cb = ctx.count_pass("primary")
    ctx.set_layout("compact")
    assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8))
    for k in ("albedo", "normal", "illumination", "frame"):
        assert np.array_equal(a[k], b[k]), k
    assert ca == cb and ca["rays"] == W * H
    assert ca["t_in"] == int(a["hits"]["trips"].astype(np.int64).sum())
    # a band of rows against the oracle at full resolution: the oracle is asked for the whole frame
    # only at 1080p (about a second); at 4K it checks the identity of hit statistics instead
    if (W, H) == (1920, 1080):
        r = oracle.render(sc.oracle_world, cam, W, H)
        assert_primary_parity(a, r)
        assert np.array_equal(a["illumination"], r["illumination"])
        assert channel_diff(a["frame"], r["frame"]).max() <= 1
        assert ca == r["primary_counters"]


@pytest.mark.parametrize("size", [(1280, 720), (1920, 1080)])
def test_c1_c2_bench_frames_full_size(uvt, oracle, w1, size):
    """BASELINE configs 1 and 2 exactly as bench.py renders them: W1, camera K0, full size, every pixel against the oracle."""
    ctx, sc = w1
    W, H = size
    ctx.set_layout("compact")
    ctx.resize(W, H)
    cam = camera_k0(oracle)
    g = gpu_render(ctx, cam)
    r = oracle.render(sc.oracle_world, cam, W, H)
    assert_primary_parity(g, r)
    assert np.array_equal(g["illumination"], r["illumination"])
    assert channel_diff(g["frame"], r["frame"]).max() <= 1
    assert ctx.count_pass("primary") == r["primary_counters"] and ctx.count_pass("secondary") == r["secondary_counters"]


def test_c4_8k_frame_tiled_full_size(uvt, oracle, w1):
    """BASELINE config 4: the 7680x4320 frame of W1 / K1 rendered as 8 interleaved 16-row band partitions (what 8 ranks render),
    assembled, and every one of its 33 M pixels compared with the oracle: hit records, G-buffer, illumination, shaded frame."""
    ctx, sc = w1
    W, H, band, n_parts = 7680, 4320, 16, 8
    cam = camera_k1(uvt, oracle)
    ctx.set_layout("compact")
    r = oracle.render(sc.oracle_world, cam, W, H)
    ctx.set_partition(band, n_parts, 0)
    ctx.resize(W, H)
    rows_seen = np.zeros(H, bool)
    try:
        for part in range(n_parts):
            ctx.set_partition(band, n_parts, part)
            g = gpu_render(ctx, cam)
            gy = uvt.tiles.local_to_global_rows(H, band, n_parts, part, g["frame"].shape[0])
            ok = gy >= 0
            rows = gy[ok]
            rows_seen[rows] = True
            part_ref = {k: r[k][rows] for k in ("albedo", "normal", "position", "hits")}
            assert_primary_parity({k: g[k][ok] for k in ("albedo", "normal", "position", "hits")}, part_ref)
            assert np.array_equal(g["illumination"][ok], r["illumination"][rows])
            assert channel_diff(g["frame"][ok], r["frame"][rows]).max() <= 1
    finally:
        ctx.set_partition(8, 1, 0)
        ctx.resize(64, 64)
    assert rows_seen.all()


# ---- W4 (procgen 2048) and the sealed-ray / free-trip machinery ----------------------------------------
@pytest.fixture(scope="module")
def w4(uvt, scene_factory):
    ctx = uvt.Context(0, hit_buffer=True, map_dim=2048)
    sc = scene_factory(2048, "procgen", ctx=ctx)
    yield ctx, sc
    ctx.close()


def test_w4_world_parity(uvt, oracle, w4):
    """BASELINE config 3 world (procgen 2048, trees in the x,z<500 corner) at a size the oracle finishes in seconds."""
    ctx, sc = w4
    assert ctx.effective_layout() == "compact"
    prm = oracle.params(2048)
    for cam in (uvt.scenes.camera_k1(2048), uvt.scenes.camera_k0(2048)):
        ctx.resize(384, 216)
        g = gpu_render(ctx, cam)
        r = oracle.render(sc.oracle_world, cam, 384, 216, prm)
        assert_primary_parity(g, r)
        assert np.array_equal(g["illumination"], r["illumination"])
        assert channel_diff(g["frame"], r["frame"]).max() <= 1
        assert ctx.count_pass("primary") == r["primary_counters"]
    # a few poses of the C5 sweep
    ctx.resize(160, 90)
    for cam in uvt.scenes.sweep_poses(2048, 6):
        g = gpu_render(ctx, cam)
        assert_primary_parity(g, oracle.render(sc.oracle_world, cam, 160, 90, prm))


def test_c3_4k_frame_full_size(uvt, oracle, w4):
    """BASELINE config 3 exactly as bench.py renders it: W4, camera K1, 3840x2160, every pixel of the primary, shadow and
    shade passes against the oracle (the oracle needs a few seconds for the 8.3 M + 5 M rays)."""
    ctx, sc = w4
    W, H = 3840, 2160
    cam = uvt.scenes.camera_k1(2048)
    prm = oracle.params(2048)
    ctx.resize(W, H)
    g = gpu_render(ctx, cam)
    r = oracle.render(sc.oracle_world, cam, W, H, prm)
    assert_primary_parity(g, r)
    assert np.array_equal(g["illumination"], r["illumination"])
    assert channel_diff(g["frame"], r["frame"]).max() <= 1
    assert ctx.count_pass("primary") == r["primary_counters"] and ctx.count_pass("secondary") == r["secondary_counters"]
    ctx.resize(64, 64)


def test_c5_pose_sweep_full_size(uvt, oracle, w4):
    """BASELINE config 5: the 256 LCG-seeded poses over W4 at 1920x1080.  Every pose is rendered on the GPU; 12 of them are
    compared with the oracle pixel for pixel, all 256 on 600 random pixels each (153,600 rays: primary record + shadow texel)."""
    ctx, sc = w4
    W, H = 1920, 1080
    prm = oracle.params(2048)
    poses = uvt.scenes.sweep_poses(2048, 256)
    ctx.resize(W, H)
    rng = np.random.default_rng(55)
    n_hits = 0
    for i, cam in enumerate(poses):
        g = gpu_render(ctx, cam)
        if i % 22 == 3:
            r = oracle.render(sc.oracle_world, cam, W, H, prm)
            assert_primary_parity(g, r)
            assert np.array_equal(g["illumination"], r["illumination"])
            assert channel_diff(g["frame"], r["frame"]).max() <= 1
        xs, ys = rng.integers(0, W, 600), rng.integers(0, H, 600)
        p = oracle.primary_pixels(sc.oracle_world, cam, W, H, xs, ys, prm)
        gh = g["hits"][ys, xs]
        for f in ("px", "py", "pz", "block", "color", "face", "trips", "exit_kind"):
            assert np.array_equal(gh[f], p["hits"][f]), (i, f)
        hit = p["hits"]["face"] != 0
        n_hits += int(hit.sum())
        assert ulp_diff(gh["distance"][hit], p["hits"]["distance"][hit]).max(initial=0) <= 2
        assert np.array_equal(g["normal"][ys, xs], p["normal"])
        assert np.array_equal(g["position"][ys, xs].view(np.uint32), p["position"].view(np.uint32))
        s = oracle.secondary(sc.oracle_world, p["normal"][None, :], p["position"][None, :, :], prm)
        assert np.array_equal(g["illumination"][ys, xs], s["illumination"][0]), i
    assert n_hits > 40000
    ctx.resize(64, 64)


def _line_hugging_world(uvt, oracle, rng, cam, W, H, n_rays, kinds):
    """Blocks placed 1-2 blocks beside / below the lines of random rays of the frame, at random distances along them: thin
    pillars up to just under the line, slabs and overhangs next to it.  The sealed-ray and free-run proofs of
    line_free_trips() must hold against geometry that hugs the rays they are about to skip."""
    cells = []
    for _ in range(n_rays):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        o, d, s = oracle.primary_ray(cam, W, H, x, y, 512)
        t = float(rng.uniform(8.0, 185.0))
        p = np.asarray(s, np.float64) + np.asarray(d, np.float64) * t / max(float(np.abs(d).sum()), 1e-6)  # about t trips along the ray
        kind = kinds[int(rng.integers(0, len(kinds)))]
        off = rng.integers(1, 3)  # 1 or 2 blocks off the line
        bx, by, bz = int(np.floor(p[0])), int(np.floor(p[1])), int(np.floor(p[2]))
        if kind == "pillar":      # a column from the ground to `off` blocks under the line
            cells += [(bx, yy, bz) for yy in range(max(by - off - int(rng.integers(0, 12)), 0), by - off + 1)]
        elif kind == "beside":    # a block at the line's height, `off` blocks to the side
            ax = int(rng.integers(0, 2))
            cells.append((bx + (off if ax == 0 else 0) * (1 if rng.random() < 0.5 else -1), by, bz + (off if ax == 1 else 0) * (1 if rng.random() < 0.5 else -1)))
        elif kind == "slab":      # a 3x3 slab `off` blocks under the line
            cells += [(bx + i, by - off, bz + j) for i in range(-1, 2) for j in range(-1, 2)]
        elif kind == "overhang":  # a slab `off` blocks ABOVE the line (the column tops then lie above the ray)
            cells += [(bx + i, by + off, bz + j) for i in range(-1, 2) for j in range(-1, 2)]
        elif kind == "on":        # a block the ray does meet
            cells.append((bx, by, bz))
    return [(x, y, z) for (x, y, z) in cells if 0 <= x < 512 and 0 <= y < 512 and 0 <= z < 512]


@pytest.mark.parametrize("seed", range(6))
def test_line_walk_against_ray_hugging_geometry(uvt, oracle, scene_factory, seed):
    """Adversarial worlds for the column-tops walk (sealed rays, long free runs): geometry placed one or two blocks off the
    lines of the very rays being traced — climbing, level and descending, lattice-aligned and generic camera positions — with
    and without a ground plane.  Fast path == verbatim-trips kernel (reference layout) == oracle, primary and shadow pass."""
    rng = np.random.default_rng(1000 + seed)
    W, H = 192, 108
    V = uvt.voxel.Voxel
    lattice = seed % 2 == 0
    pos = (256.0, 40.0 + 8 * seed, 256.0) if lattice else (float(rng.uniform(200, 312)), float(rng.uniform(20, 120)), float(rng.uniform(200, 312)))
    pitch = [-0.5, -0.15, 0.0, 0.12, 0.35, 0.7][seed]
    cam = oracle.make_camera(pos, pitch_yaw_matrix(uvt, pitch, float(rng.uniform(0, 2 * np.pi))), fov=float(rng.uniform(0.9, 1.9)))
    kinds = [["pillar", "beside", "slab", "on"], ["pillar", "beside", "slab", "overhang", "on"]][seed % 2]

    def fill(bm):
        if seed % 3 != 2:   # a ground plane 20-60 blocks under the camera, so that descending rays have something to approach
            gy = max(int(pos[1]) - int(rng.integers(20, 60)), 0)
            for x in range(176, 336, 1):
                for z in range(176, 336, 1):
                    bm.set(x, gy, z, V(21 + (x + z) % 3, True))
        for (x, y, z) in _line_hugging_world(uvt, oracle, rng, cam, W, H, 140, kinds):
            bm.set(x, y, z, V(int(rng.integers(0, 29)), True))

    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(512, fill, ctx=ctx)
        ctx.resize(W, H)
        ctx.set_layout("compact")
        a = gpu_render(ctx, cam)
        ctx.set_layout("reference")
        b = gpu_render(ctx, cam)
        ctx.set_layout("compact")
        assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8))
        for k in ("albedo", "normal", "illumination", "frame"):
            assert np.array_equal(a[k], b[k]), k
        r = oracle.render(sc.oracle_world, cam, W, H)
        assert_primary_parity(a, r)
        assert np.array_equal(a["illumination"], r["illumination"])
        stats, exact = ctx.fetch_stats("primary"), ctx.count_pass("primary")
        assert stats["lookups"] < exact["t_in"]   # the walk and the clearances did skip fetches in this world


def test_line_walk_w4_near_map_faces_and_lattice_origins(uvt, oracle, w4):
    """W4 cameras a few blocks from the map faces and on lattice points, looking out, along and into the map, up and down:
    the walk must stop proving at the faces (a ray that leaves the map is an exit, not an iteration-cap miss)."""
    ctx, sc = w4
    prm = oracle.params(2048)
    ctx.resize(160, 90)
    rng = np.random.default_rng(77)
    cams = []
    for (x, z) in ((3.0, 1000.0), (2044.5, 700.25), (900.0, 2.0), (1200.0, 2045.0), (4.0, 4.0), (2040.0, 2040.0), (1024.0, 1024.0)):
        h = max(uvt.procgen.height(2048, int(x), int(z)), 16)
        for k in range(3):
            cams.append(oracle.make_camera((x, float(h + rng.integers(2, 60)), z),
                                           pitch_yaw_matrix(uvt, float(rng.uniform(-0.9, 0.9)), float(rng.uniform(0, 2 * np.pi))), fov=float(rng.uniform(0.7, 2.0))))
    cams.append(oracle.make_camera((1024.0, 2040.0, 1024.0), pitch_yaw_matrix(uvt, -0.4, 1.0)))   # just under the top face, looking up
    cams.append(oracle.make_camera((1024.0, 2040.0, 1024.0), pitch_yaw_matrix(uvt, 1.2, 1.0)))    # ... and down
    n_exit = 0
    for cam in cams:
        g = gpu_render(ctx, cam)
        r = oracle.render(sc.oracle_world, cam, 160, 90, prm)
        assert_primary_parity(g, r)
        assert np.array_equal(g["illumination"], r["illumination"])
        n_exit += int((r["hits"]["exit_kind"] == 2).sum())
    assert n_exit > 20000   # the faces were exercised


def test_sealed_rays_and_map_faces(uvt, oracle, w1):
    """Rays that climb above every occupied block are retired early (sealed) — results must still equal the
    192-trip march; cameras near a map face must NOT seal (the ray may leave the map first)."""
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(160, 90)
    up = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float32)  # looking straight up
    cams = [oracle.make_camera((256.0, 30.0, 256.0), up),                                     # sealed almost at once
            oracle.make_camera((256.0, 200.0, 256.0), pitch_yaw_matrix(uvt, -0.5, 1.0)),      # high, looking up
            oracle.make_camera((256.0, 500.0, 256.0), up),                                     # leaves through the top face
            oracle.make_camera((8.0, 60.0, 8.0), pitch_yaw_matrix(uvt, -0.3, 3.9)),           # corner, looking out and up
            oracle.make_camera((500.0, 40.0, 256.0), pitch_yaw_matrix(uvt, -0.2, np.pi / 2)),  # near +x face, looking out
            oracle.make_camera((256.0, 25.0, 256.0), pitch_yaw_matrix(uvt, -0.9, 0.3))]
    for cam in cams:
        g = gpu_render(ctx, cam)
        r = oracle.render(sc.oracle_world, cam, 160, 90)
        assert_primary_parity(g, r)
        assert np.array_equal(g["illumination"], r["illumination"])
    # odd step caps: sealing and the lockstep cap must follow maxSteps
    for cap in (1, 7, 64, 500):
        ctx.set_max_steps(cap, 48)
        g = gpu_render(ctx, camera_k1(uvt, oracle))
        r = oracle.render(sc.oracle_world, camera_k1(uvt, oracle), 160, 90, oracle.params(512, primary_max_steps=cap))
        assert_primary_parity(g, r)
    ctx.set_max_steps(192, 48)


def test_fast_path_fetches_fewer_trips(uvt, oracle, w1):
    """The B200 layout must actually skip fetches: lookups performed < trips executed (and counters stay exact)."""
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(320, 180)
    ctx.set_camera(camera_k0(oracle))
    exact = ctx.count_pass("primary")
    stats = ctx.fetch_stats("primary")
    assert stats["rays"] == exact["rays"] == 320 * 180 and stats["hits"] == exact["hits"]
    assert stats["lookups"] < 0.6 * exact["t_in"]


def test_pipelined_readback_matches_blocking(uvt, oracle, w1):
    """uvt_readback_async snapshots the buffer on the ctx stream, so the next dispatch may overwrite it at once."""
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(320, 180)
    cams = [camera_k0(oracle), camera_k1(uvt, oracle), oracle.make_camera((300.0, 40.0, 200.0), pitch_yaw_matrix(uvt, 0.3, 4.0))]
    expect = []
    for cam in cams:
        ctx.set_camera(cam); ctx.dispatch_primary()
        expect.append(ctx.readback("albedo").copy())
    bufs = [ctx.pinned_empty(320 * 180 * 4, np.uint32) for _ in cams]
    for cam, buf in zip(cams, bufs):          # three frames in flight through two snapshot slots
        ctx.set_camera(cam); ctx.dispatch_primary()
        ctx.readback_async("albedo", buf)
    ctx.readback_wait()
    for buf, e in zip(bufs, expect):
        assert np.array_equal(buf.reshape(180, 320), e)


def test_fast_path_vs_verbatim_kernel_many_poses(uvt, oracle, w1):
    """Free trips, sealed rays and the dense grid against the verbatim-trips kernel on the reference layout (itself held
    to the oracle above), GPU vs GPU, over many random poses: looking up, down, along the horizon, near map faces,
    narrow and wide fields of view, primary and shadow pass."""
    ctx, sc = w1
    ctx.resize(256, 144)
    rng = np.random.default_rng(2026)
    for i in range(60):
        near_face = i % 6 == 0
        x, z = (rng.uniform(2, 30, 2) if near_face else rng.uniform(20, 490, 2))
        if near_face and i % 12 == 0:
            x, z = 512 - x, 512 - z
        y = uvt.procgen.height(512, int(x), int(z)) + rng.uniform(1.5, 60 if i % 5 else 300)
        pitch = rng.uniform(-1.5, 1.5) if i % 4 == 0 else rng.uniform(-0.5, 0.5)
        cam = oracle.make_camera((x, y, z), pitch_yaw_matrix(uvt, pitch, rng.uniform(0, 2 * np.pi)), fov=rng.uniform(0.4, 2.3))
        ctx.set_layout("compact")
        a = gpu_render(ctx, cam)
        ctx.set_layout("reference")
        b = gpu_render(ctx, cam)
        assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8)), f"pose {i}"
        for k in ("albedo", "normal", "illumination", "frame"):
            assert np.array_equal(a[k], b[k]), (i, k)
    ctx.set_layout("compact")


def test_lattice_and_diagonal_rays_round_up_carries(uvt, oracle, w1):
    """Cameras on lattice points looking along face and space diagonals: many rays cross block edges and corners exactly,
    `within` rounds up to the step size (the carry of the free-trip bound) and t ties are common.  The conditional last
    free trip must fall back to the generic loop there: fast path == verbatim kernel == oracle."""
    ctx, sc = w1
    ctx.resize(256, 144)
    n_checked = 0
    for (pos, pitch, yaw, fov) in [((256.0, 40.0, 256.0), 0.0, np.pi / 4, np.pi / 2),
                                   ((256.0, 40.0, 256.0), 0.0, 3 * np.pi / 4, np.pi / 2),
                                   ((128.0, 64.0, 128.0), float(np.arctan(1 / np.sqrt(2))), np.pi / 4, 1.2),
                                   ((128.0, 64.0, 128.0), -float(np.arctan(1 / np.sqrt(2))), 5 * np.pi / 4, 1.2),
                                   ((300.0, 30.0, 200.0), 0.0, 0.0, np.pi / 2),
                                   ((300.5, 30.5, 200.5), 0.0, np.pi / 2, 2 * float(np.arctan(0.5))),
                                   ((64.0, 80.0, 64.0), -np.pi / 4, np.pi / 4, np.pi / 2),
                                   ((8.0, 24.0, 8.0), 0.25, np.pi / 4, np.pi / 2)]:
        cam = oracle.make_camera(pos, pitch_yaw_matrix(uvt, pitch, yaw), fov=fov)
        ctx.set_layout("compact")
        a = gpu_render(ctx, cam)
        ctx.set_layout("reference")
        b = gpu_render(ctx, cam)
        ctx.set_layout("compact")
        assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8)), (pos, pitch, yaw)
        for k in ("albedo", "normal", "illumination", "frame"):
            assert np.array_equal(a[k], b[k]), (pos, k)
        assert_primary_parity(a, oracle.render(sc.oracle_world, cam, 256, 144))
        n_checked += 1
    assert n_checked == 8


def test_degenerate_camera_rays_take_the_generic_loop_on_the_compact_layout(uvt, oracle, w1):
    """A camera matrix that squeezes the x axis to 1e-36 gives every ray a non-finite-ish reciprocal (|1/dir.x| > 1e30):
    those lanes leave the fast path for the generic loop, which must read the compact layout (clearance bytes of empty
    blocks are not materials) exactly like the reference layout."""
    ctx, sc = w1
    ctx.resize(128, 72)
    m = pitch_yaw_matrix(uvt, 0.3, 0.7).reshape(4, 4).copy()
    m[0, :] *= np.float32(1e-36)
    for pos in ((256.0, 40.0, 256.0), (100.5, 30.25, 300.75)):
        cam = oracle.make_camera(pos, m.reshape(16))
        ctx.set_layout("compact")
        a = gpu_render(ctx, cam)
        ctx.set_layout("reference")
        b = gpu_render(ctx, cam)
        ctx.set_layout("compact")
        assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8))
        for k in ("albedo", "normal", "illumination", "frame"):
            assert np.array_equal(a[k], b[k]), k
        assert_primary_parity(a, oracle.render(sc.oracle_world, cam, 128, 72))
        assert (a["hits"]["exit_kind"] == 0).any()   # the view does reach terrain


# ---- incremental publish: uvt_world_commit_region / uvt_world_set_voxel / bind() of a dirty box ------
def _fresh_full_commit(uvt, sc, tmp_path, **ctx_kw):
    """A second ctx holding the same host world, published by one full commit."""
    path = str(tmp_path / "world.uvtw")
    sc.bm.save(path)
    ctx = uvt.Context(0, hit_buffer=True, **ctx_kw)
    bm = uvt.voxel.VoxelBrickmap.load(path, ctx)
    atlas = uvt.voxel.VoxelModelAtlas.init(ctx)
    for m in sc.models:
        atlas.append_model(m)
    bm.bind(9)
    return ctx, bm


def _same_render(a, b):
    assert np.array_equal(a["hits"].view(np.uint8), b["hits"].view(np.uint8))
    for k in ("albedo", "normal", "position", "illumination", "frame"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("dense", [True, False])
def test_incremental_commit_equals_full_commit(uvt, oracle, scene_factory, tmp_path, dense):
    """Edits published through bind() (dirty box -> uvt_world_commit_region) leave the same derived layout
    (checksums) and the same pixels as a fresh full commit of the edited world; hits match the oracle."""
    V = uvt.voxel.Voxel
    rng = np.random.default_rng(11)
    dim = 128
    with uvt.Context(0, hit_buffer=True, dense=dense) as ctx:
        sc = scene_factory(dim, "procgen", ctx=ctx)
        ctx.resize(160, 96)
        cams = [oracle.make_camera((64.0, 30.0, 40.0)),
                oracle.make_camera((30.0, 45.0, 30.0), pitch_yaw_matrix(uvt, 0.5, 0.8)),
                oracle.make_camera((100.0, 60.0, 100.0), pitch_yaw_matrix(uvt, 0.9, 3.9))]

        def edits(step):
            if step == 0:    # dig: remove blocks inside existing bricks
                for _ in range(40):
                    x, z = (int(v) for v in rng.integers(40, 90, 2))
                    for y in range(10, 24):
                        if sc.bm.get(x, y, z):
                            sc.bm.set(x, y, z, 0)
            elif step == 1:  # build inside existing bricks and just above them
                for _ in range(30):
                    x, z = (int(v) for v in rng.integers(50, 80, 2))
                    sc.bm.set(x, int(rng.integers(16, 24)), z, V(11, True))
            elif step == 2:  # a floating slab in empty air: new bricks, new virtual bricks, new chunk distances
                for x in range(60, 71):
                    for z in range(44, 52):
                        sc.bm.set(x, 52, z, V(13, True))
            elif step == 3:  # single block at the map corner and one at the top face
                sc.bm.set(0, 40, 0, V(11, True))
                sc.bm.set(dim - 1, dim - 1, dim - 1, V(11, True))
            elif step == 4:  # remove part of the slab again (bricks stay allocated, chunks stay non-empty)
                for x in range(60, 66):
                    for z in range(44, 52):
                        sc.bm.set(x, 52, z, 0)
            elif step == 5:  # a tall pillar crossing many chunk rows
                for y in range(0, 100):
                    sc.bm.set(20, y, 90, V(21, True))

        for step in range(6):
            edits(step)
            launches0 = ctx.launch_count()
            sc.bm.bind(9)
            n_launch = ctx.launch_count() - launches0
            assert 0 < n_launch < 20, n_launch
            ctx2, bm2 = _fresh_full_commit(uvt, sc, tmp_path, dense=dense)
            try:
                assert ctx.world_layout_checksum()[:4] == ctx2.world_layout_checksum()[:4], step
                ctx2.resize(160, 96)
                world = oracle.World(dim, sc.bm.chunks().copy(), sc.bm.bricks().copy(), sc.oracle_world.atlas)
                for cam in cams:
                    a, b = gpu_render(ctx, cam), gpu_render(ctx2, cam)
                    _same_render(a, b)
                    assert ctx.count_pass("primary") == ctx2.count_pass("primary")
                    assert_primary_parity(a, oracle.render(world, cam, 160, 96))
            finally:
                ctx2.close()
        # a clean map: bind() is free
        launches0 = ctx.launch_count()
        sc.bm.bind(9)
        assert ctx.launch_count() == launches0


def test_incremental_commit_new_material_and_set_voxel(uvt, oracle, scene_factory, tmp_path):
    """A block word never seen before gets a material id in place; uvt_world_set_voxel follows map_setVoxel."""
    V = uvt.voxel.Voxel
    dim = 128
    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(dim, "procgen", ctx=ctx)
        ctx.resize(160, 96)
        cam = oracle.make_camera((64.0, 30.0, 40.0))
        before = gpu_render(ctx, cam)
        n_mats = ctx.world_layout_checksum()[4]
        for x in range(60, 68):
            for y in range(18, 30):
                sc.bm.set(x, y, 60, V(5, False))    # model 5 without the "solid" flag: a new block word
        sc.bm.bind(9)
        assert ctx.world_layout_checksum()[4] == n_mats + 1
        world = oracle.World(dim, sc.bm.chunks().copy(), sc.bm.bricks().copy(), sc.oracle_world.atlas)
        after = gpu_render(ctx, cam)
        assert_primary_parity(after, oracle.render(world, cam, 160, 96))
        assert not np.array_equal(before["hits"]["block"], after["hits"]["block"])
        ctx2, _ = _fresh_full_commit(uvt, sc, tmp_path)
        try:
            ctx2.resize(160, 96)
            _same_render(after, gpu_render(ctx2, cam))
        finally:
            ctx2.close()
        # map_setVoxel: lands in an existing brick, ignored in an empty chunk and outside the map
        assert ctx.world_set_voxel(64, 15, 45, V(11, True)) is True
        assert ctx.world_set_voxel(64, 120, 45, V(11, True)) is False
        assert ctx.world_set_voxel(dim, 20, 45, V(11, True)) is False
        assert sc.bm.get(64, 15, 45) == V(11, True) and sc.bm.get(64, 120, 45) == 0
        world = oracle.World(dim, sc.bm.chunks().copy(), sc.bm.bricks().copy(), sc.oracle_world.atlas)
        assert_primary_parity(gpu_render(ctx, cam), oracle.render(world, cam, 160, 96))


def test_incremental_commit_on_w1_near_camera(uvt, oracle, scene_factory):
    """W1 at K0: a wall built block by block, one bind per block (the interactive editing pattern)."""
    V = uvt.voxel.Voxel
    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(512, "procgen", ctx=ctx)
        ctx.resize(160, 90)
        cam = camera_k0(oracle)
        for i, (x, y) in enumerate((x, y) for y in range(22, 34) for x in range(252, 260)):
            sc.bm.set(x, y, 266, V(11, True))
            sc.bm.bind(9)
            if i % 24 == 23:
                world = oracle.World(512, sc.bm.chunks().copy(), sc.bm.bricks().copy(), sc.oracle_world.atlas)
                assert_primary_parity(gpu_render(ctx, cam), oracle.render(world, cam, 160, 90))
        ctx.set_layout("reference")
        ref = gpu_render(ctx, cam)
        ctx.set_layout("compact")
        _same_render(gpu_render(ctx, cam), ref)


# ---- several GPUs in one process (uvt_group): members may share a device, so one GPU is enough to test the plumbing ----
def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_group_frame_equals_single_ctx(uvt, oracle, scene_factory, devices):
    cam = camera_k1(uvt, oracle)
    with uvt.Context(0) as ctx, uvt.Group(devices) as grp:
        sc1 = scene_factory(512, "procgen", ctx=ctx)
        scene_factory(512, "procgen", ctx=grp)
        assert grp.size == len(devices)
        for (W, H) in ((320, 200), (1920, 1080)):   # 200 rows: a ragged last band
            ctx.resize(W, H)
            ctx.set_camera(cam)
            ctx.dispatch_frame()
            ref = ctx.readback("frame")
            grp.resize(W, H)
            grp.set_camera(cam)
            grp.dispatch_frame()
            got = grp.readback_frame()
            assert np.array_equal(got, ref)
            for which in ("primary", "secondary"):
                assert grp.count_pass(which) == ctx.count_pass(which)
        r = oracle.render(sc1.oracle_world, cam, 320, 200)
        grp.resize(320, 200)
        grp.dispatch_frame()
        assert channel_diff(grp.readback_frame(), r["frame"]).max() <= 1


def test_group_world_edits_reach_every_member(uvt, oracle, scene_factory):
    V = uvt.voxel.Voxel
    cam = camera_k0(oracle)
    with uvt.Context(0) as ctx, uvt.Group([0, 0]) as grp:
        a = scene_factory(512, "procgen", ctx=ctx)
        b = scene_factory(512, "procgen", ctx=grp)
        for c in (ctx, grp):
            c.resize(320, 192)
            c.set_camera(cam)
        for step in range(3):
            for sc in (a, b):
                for y in range(20, 30 + 4 * step):
                    for x in range(250, 262):
                        sc.bm.set(x, y, 268 + step, V(11, True))    # walls in front of the camera: new bricks, both bands
                sc.bm.bind(9)
            ctx.dispatch_frame()
            grp.dispatch_frame()
            assert np.array_equal(grp.readback_frame(), ctx.readback("frame")), step


def test_group_across_two_gpus(uvt, oracle, scene_factory):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cam = camera_k1(uvt, oracle)
    with uvt.Context(0) as ctx, uvt.Group(list(range(_n_gpus()))) as grp:
        scene_factory(512, "procgen", ctx=ctx)
        scene_factory(512, "procgen", ctx=grp)
        for c in (ctx, grp):
            c.resize(1920, 1080)
            c.set_camera(cam)
        ctx.dispatch_frame()
        grp.dispatch_frame()
        assert np.array_equal(grp.readback_frame(), ctx.readback("frame"))


def test_incremental_commit_random_edit_sequences(uvt, oracle, scene_factory, tmp_path):
    """Forty random edit batches on a 64^3 world (single blocks, boxes, columns; set and clear; map faces and corners),
    one bind() each: after every batch the derived layout equals a fresh full commit's, and every fifth batch the
    pixels do too."""
    V = uvt.voxel.Voxel
    rng = np.random.default_rng(2024)
    dim = 64
    mats = [V(11, True), V(13, True), V(21, True), V(8, False), 0, 0]
    with uvt.Context(0, hit_buffer=True) as ctx:
        sc = scene_factory(dim, "procgen", ctx=ctx)
        ctx.resize(96, 64)
        cam = oracle.make_camera((32.0, 30.0, 4.0), pitch_yaw_matrix(uvt, 0.45, 0.0))
        for batch in range(40):
            kind = int(rng.integers(0, 4))
            if kind == 0:      # scattered single blocks, anywhere (faces and corners included)
                for _ in range(int(rng.integers(1, 6))):
                    x, y, z = (int(v) for v in rng.choice([0, 1, 7, 8, 31, 32, 55, 56, 62, 63], 3))
                    sc.bm.set(x, y, z, mats[int(rng.integers(0, len(mats)))])
            elif kind == 1:    # a small box
                lo = rng.integers(0, dim - 6, 3)
                ext = rng.integers(1, 6, 3)
                m = mats[int(rng.integers(0, len(mats)))]
                for x in range(lo[0], lo[0] + ext[0]):
                    for y in range(lo[1], lo[1] + ext[1]):
                        for z in range(lo[2], lo[2] + ext[2]):
                            sc.bm.set(int(x), int(y), int(z), m)
            elif kind == 2:    # a column through every chunk row
                x, z = (int(v) for v in rng.integers(0, dim, 2))
                m = mats[int(rng.integers(0, len(mats)))]
                for y in range(int(rng.integers(0, 8)), dim, int(rng.integers(1, 4))):
                    sc.bm.set(x, y, z, m)
            else:              # dig around the terrain surface
                for _ in range(20):
                    x, z = (int(v) for v in rng.integers(0, dim, 2))
                    sc.bm.set(x, int(rng.integers(0, 20)), z, 0)
            sc.bm.bind(9)
            ctx2, _ = _fresh_full_commit(uvt, sc, tmp_path)
            try:
                assert ctx.world_layout_checksum()[:4] == ctx2.world_layout_checksum()[:4], (batch, kind)
                if batch % 5 == 4:
                    ctx2.resize(96, 64)
                    a = gpu_render(ctx, cam)
                    _same_render(a, gpu_render(ctx2, cam))
                    world = oracle.World(dim, sc.bm.chunks().copy(), sc.bm.bricks().copy(), sc.oracle_world.atlas)
                    assert_primary_parity(a, oracle.render(world, cam, 96, 64))
            finally:
                ctx2.close()


def test_caller_owned_staging(uvt, oracle, world64, models):
    """uvt_world_use_staging: the ctx publishes a world that lives in the caller's (pageable) arrays; growing the pool
    on the caller's side is announced with dim = 0."""
    import ctypes
    L = uvt._native.load()
    with uvt.Context(0, map_dim=64, hit_buffer=True) as ctx:
        atlas = uvt.voxel.VoxelModelAtlas.init(ctx)
        for m in models:
            atlas.append_model(m)
        chunks = world64.chunks.copy()
        bricks = world64.bricks.copy()
        ctx.check(L.uvt_world_use_staging(ctx.handle, 64, chunks.ctypes.data, bricks.ctypes.data, bricks.shape[0]))
        assert L.uvt_world_grow(ctx.handle, 4 * bricks.shape[0], ctypes.byref(ctypes.c_void_p())) == uvt._native.UVT_ERR_INVALID
        ctx.check(L.uvt_world_commit(ctx.handle, world64.n_bricks))
        ctx.resize(96, 64)
        cam = oracle.make_camera((35.5, 20.0, 4.0))
        a = gpu_render(ctx, cam)
        assert_primary_parity(a, oracle.render(world64.oracle_world, cam, 96, 64))
        # the caller grows its pool, adds a brick of rock in an empty chunk and publishes only that box
        bigger = np.zeros((bricks.shape[0] + 8, 512), dtype=np.uint32)
        bigger[:bricks.shape[0]] = bricks
        n = world64.n_bricks
        cz, cy, cx = 1, 2, 4                      # chunk (4, 2, 1): just above the water of a 64^3 world, straight ahead of the camera
        assert chunks.reshape(8, 8, 8)[cz, cy, cx] == 0
        chunks.reshape(8, 8, 8)[cz, cy, cx] = n + 1
        bigger[n, :] = uvt.voxel.Voxel(11, True)
        ctx.check(L.uvt_world_use_staging(ctx.handle, 0, None, bigger.ctypes.data, bigger.shape[0]))
        lo = (ctypes.c_uint32 * 3)(32, 16, 8)
        hi = (ctypes.c_uint32 * 3)(39, 23, 15)
        ctx.check(L.uvt_world_commit_region(ctx.handle, n + 1, ctypes.byref(lo), ctypes.byref(hi)))
        b = gpu_render(ctx, cam)
        world = oracle.World(64, chunks.copy(), bigger[:n + 1].copy(), world64.oracle_world.atlas)
        assert_primary_parity(b, oracle.render(world, cam, 96, 64))
        assert not np.array_equal(a["hits"]["block"], b["hits"]["block"])


# ---- procgen on the device ---------------------------------------------------------------------------
@pytest.mark.parametrize("dim,off", [(64, (0.0, 0.0)), (128, (0.0, 0.0)), (512, (0.0, 0.0)), (512, (1234.5, -77.25)), (1024, (300.0, 900.0)), (2048, (0.0, 0.0))])
def test_device_procgen_equals_host_procgen(uvt, dim, off):
    """uvt_procgen_device: the world of the serial host procgen (src/procgen.zig restated in csrc/host/world.cpp) byte for byte —
    chunk table, brick numbering (first-touch order of the allocator), brick words, pool capacity — for several sizes and noise
    offsets (other offsets move the hills, the water line crossings and the trees)."""
    host = uvt.voxel.VoxelBrickmap.init(dim, 8, None)
    uvt.procgen.procgen(dim, host, *off, device="host")
    with uvt.Context(0, map_dim=dim) as ctx:
        dev = uvt.voxel.VoxelBrickmap.init(dim, 8, ctx)
        uvt.procgen.procgen(dim, dev, *off, device="device")
        assert dev.n_bricks == host.n_bricks and dev.capacity == host.capacity
        assert np.array_equal(dev.chunks(), host.chunks())
        hb, db = host.bricks(), dev.bricks()
        if not np.array_equal(db, hb):
            bad = np.argwhere(db.reshape(-1, 512) != hb.reshape(-1, 512))
            raise AssertionError(f"{len(bad)} brick words differ, first at brick/offset {bad[:5].tolist()}")
        dev.bind(9)   # and it commits like any other world
        assert ctx.effective_layout() == "compact"


def test_device_procgen_is_the_default_for_a_fresh_map_on_a_ctx(uvt, oracle, models):
    """procgen(..., device="auto"): device path for an empty map attached to a ctx, host path otherwise; same frame either way."""
    cam = uvt.scenes.camera_k1(512)
    frames = []
    for mode in ("auto", "host"):
        with uvt.Context(0) as ctx:
            bm = uvt.voxel.VoxelBrickmap.init(512, 8, ctx)
            uvt.procgen.procgen(512, bm, device=mode)
            atlas = uvt.voxel.VoxelModelAtlas.init(ctx)
            for m in models:
                atlas.append_model(m)
            bm.bind(9)
            ctx.resize(320, 180)
            ctx.set_camera(cam)
            ctx.dispatch_frame()
            frames.append(ctx.readback("frame").copy())
    assert np.array_equal(frames[0], frames[1])


# ---- entities done properly: map.glsl:203-248 live + the primary composite of primary.comp.glsl:45-54 ----
ENTITY_CAMERAS = [((258.0, 25.0, 262.0), 0.5, 3.6), ((262.0, 30.0, 262.0), 0.6, 5 * np.pi / 4), ((249.0, 27.0, 262.0), 0.9, 2.2),
                  ((254.5, 21.6, 250.0), 0.05, 0.0), ((256.5, 21.5, 256.5), 0.2, 1.0)]  # the last one sits INSIDE entity 0's box


def assert_frame_parity(g, r):
    assert_primary_parity(g, r)
    assert np.array_equal(g["illumination"], r["illumination"])
    assert channel_diff(g["frame"], r["frame"]).max() <= 1


@pytest.mark.parametrize("case", ["atlas8", "chicken32", "boxes-custom"])
def test_entity_models_frame_parity(uvt, oracle, w1, case):
    import os
    from conftest import GOLDEN
    ctx, sc = w1
    ctx.set_layout("compact")
    W, H = 256, 144
    try:
        if case == "atlas8":      # the text as written: five literal boxes, model = texels [0,8)^3 of the atlas
            ctx.set_entity_mode("models")
            ent = oracle.entities("models")
        elif case == "chicken32":  # chicken.vox (game.zig:114) as a 32^3 model in 4-block boxes
            m = np.load(os.path.join(GOLDEN, "chicken_32.npy"))
            pos = [(252.0, 21.0, 254.0), (258.5, 22.25, 259.0), (246.0, 20.0, 262.0)]
            ctx.set_entity_mode("models")
            ctx.set_entities(pos)
            ctx.entity_model_upload(m, 32, 256)
            ent = oracle.entities("models", positions=pos, model=m, size=32, max_steps=256)
        else:                      # live behaviour (boxes as lines, shadow pass only) with a custom entity list
            pos = [(255.0, 22.0, 258.0), (259.0, 21.0, 255.0), (250.0, 23.0, 250.0), (262.0, 21.0, 262.0)]
            ctx.set_entities(pos)
            ent = oracle.entities("boxes", positions=pos)
        ctx.resize(W, H)
        prm = oracle.params(512, ent=ent)
        seen = 0
        for p, pitch, yaw in ENTITY_CAMERAS:
            cam = oracle.make_camera(p, pitch_yaw_matrix(uvt, pitch, yaw))
            r = oracle.render(sc.oracle_world, cam, W, H, prm)
            off = oracle.render(sc.oracle_world, cam, W, H, oracle.params(512, entities=False))
            seen += int((r["hits"]["exit_kind"] == 3).sum()) + int((r["illumination"] != off["illumination"]).sum())
            for three in (True, False):
                assert_frame_parity(gpu_render(ctx, cam, three_pass=three), r)
        assert seen > 2000, seen
    finally:
        ctx.set_entity_mode("boxes")
        ctx.set_entities(None)
        ctx.entity_model_upload(None)
    # back to the reference as it runs
    cam = oracle.make_camera(ENTITY_CAMERAS[1][0], pitch_yaw_matrix(uvt, *ENTITY_CAMERAS[1][1:]))
    assert_frame_parity(gpu_render(ctx, cam), oracle.render(sc.oracle_world, cam, W, H))


def test_entity_models_with_every_dispatch_path(uvt, oracle, scene_factory):
    """The composite and the entity shadow pass follow the frame through the pooled scheduler, the reference layout, a
    batched dispatch and a band partition; a ctx created WITHOUT the hit buffer gets one when the mode is switched."""
    p, pitch, yaw = ENTITY_CAMERAS[0]
    cam = oracle.make_camera(p, pitch_yaw_matrix(uvt, pitch, yaw))
    cam2 = oracle.make_camera(ENTITY_CAMERAS[2][0], pitch_yaw_matrix(uvt, *ENTITY_CAMERAS[2][1:]))
    W, H = 200, 120
    with uvt.Context(0, hit_buffer=False) as ctx:
        sc = scene_factory(512, "procgen", ctx=ctx)
        ctx.resize(W, H)
        ctx.set_entity_mode("models")   # after the resize: the G-buffer is re-created with a hit buffer
        prm = oracle.params(512, ent=oracle.entities("models"))
        r, r2 = (oracle.render(sc.oracle_world, c, W, H, prm) for c in (cam, cam2))
        assert (r["hits"]["exit_kind"] == 3).sum() > 300
        for layout, sched in (("compact", "tile"), ("compact", "pool"), ("reference", "tile")):
            ctx.set_layout(layout)
            ctx.set_scheduler(sched)
            assert_frame_parity(gpu_render(ctx, cam), r)
        ctx.set_layout("compact")
        ctx.set_scheduler("tile")
        # batched poses
        ctx.set_camera(np.array([cam, cam2]))
        ctx.dispatch_frame()
        for k, ref in enumerate((r, r2)):
            assert np.array_equal(ctx.readback("illumination")[k], ref["illumination"])
            assert np.array_equal(ctx.readback("hit")[k]["exit_kind"], ref["hits"]["exit_kind"])
            assert channel_diff(ctx.readback("frame")[k], ref["frame"]).max() <= 1
        # band partition: the union of two parts is the frame
        ctx.set_camera(cam)
        parts = []
        for part in range(2):
            ctx.set_partition(16, 2, part)
            ctx.resize(W, H)
            ctx.set_camera(cam)
            ctx.dispatch_frame()
            parts.append((ctx.readback("frame"), ctx.readback("hit")))
        full = np.zeros((H, W), np.uint32)
        kinds = np.zeros((H, W), np.uint8)
        for part, (f, h) in enumerate(parts):
            rows = [y for y in range(H) if (y // 16) % 2 == part]
            full[rows] = f.reshape(-1, W)[:len(rows)]
            kinds[rows] = h.reshape(-1, W)[:len(rows)]["exit_kind"]
        assert channel_diff(full, r["frame"]).max() <= 1 and np.array_equal(kinds, r["hits"]["exit_kind"])


# ---- the CUDA path against the reference's OWN shader text (oracle/_ref/libglslref.so, oracle/glsl_ref/) ------------------
def assert_gbuffer_equals_reference_text(g, ref):
    """ref: images produced by the reference's shader text.  Hits (position.w == 1) bit-exact in every image; sky albedo and
    the shaded frame within 1/255 (integer powers by multiplication vs powf)."""
    hit = ref["position"][..., 3] == 1.0
    assert np.array_equal(g["normal"], ref["normal"])
    assert np.array_equal(g["position"].view(np.uint32), ref["position"].view(np.uint32))
    assert np.array_equal(g["illumination"], ref["illumination"])
    assert np.array_equal(g["albedo"][hit], ref["albedo"][hit])
    assert channel_diff(g["albedo"], ref["albedo"]).max() <= 1
    assert channel_diff(g["frame"], ref["frame"]).max() <= 1
    return int(hit.sum())


def test_committed_reference_shader_golden(uvt, oracle, w1):
    """tests/golden/glslref_w1_*.npz: outputs of the reference's shader text (made in the build container by
    tools/make_golden.py, reproduced there by tests/test_glsl_reference.py) against the GPU."""
    import os
    from conftest import GOLDEN
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(96, 54)
    for name in ("k0_96x54", "k1_96x54"):
        gold = np.load(os.path.join(GOLDEN, f"glslref_w1_{name}.npz"))
        cam = np.frombuffer(gold["camera"].tobytes(), dtype=oracle.CAMERA_DTYPE)[0]
        assert assert_gbuffer_equals_reference_text(gpu_render(ctx, cam), gold) > 2000


@pytest.mark.parametrize("size", [(1280, 720), (1920, 1080)])
def test_gpu_equals_the_reference_shader_text_full_size(uvt, oracle, w1, size):
    """BASELINE configs 1 and 2 at full size: every G-buffer image, the illumination image and the frame of the CUDA path
    against the reference's own GLSL compiled for the CPU (the prebuilt oracle/_ref library travels to the GPU box)."""
    from oracle import glslref
    if not glslref.available():
        pytest.skip("oracle/_ref/libglslref.so is not here")
    ctx, sc = w1
    ctx.set_layout("compact")
    W, H = size
    ctx.resize(W, H)
    for cam in (camera_k0(oracle), camera_k1(uvt, oracle)):
        ref = glslref.render(sc.oracle_world, cam, W, H)
        assert assert_gbuffer_equals_reference_text(gpu_render(ctx, cam), ref) > W * H // 3


def test_gpu_equals_the_reference_shader_text_w4_and_entities(uvt, oracle, w4, w1):
    from oracle import glslref
    if not glslref.available() or not glslref.available("entities"):
        pytest.skip("oracle/_ref is not here")
    ctx, sc = w4
    W, H = 960, 540
    ctx.resize(W, H)
    for cam in (uvt.scenes.camera_k1(2048), uvt.scenes.sweep_poses(2048, 6)[5]):
        assert assert_gbuffer_equals_reference_text(gpu_render(ctx, cam), glslref.render(sc.oracle_world, cam, W, H)) > W * H // 5
    # SURVEY 8 f3 against the dead code made live (map.glsl:199 deleted, primary.comp.glsl:47-54 uncommented)
    ctx, sc = w1
    ctx.set_layout("compact")
    ctx.resize(320, 180)
    try:
        ctx.set_entity_mode("models")
        ctx.resize(320, 180)
        for p, pitch, yaw in ENTITY_CAMERAS:
            cam = oracle.make_camera(p, pitch_yaw_matrix(uvt, pitch, yaw))
            assert_gbuffer_equals_reference_text(gpu_render(ctx, cam), glslref.render(sc.oracle_world, cam, 320, 180, variant="entities"))
    finally:
        ctx.set_entity_mode("boxes")


# ---- the per-column sun clearance map (sun1): shadow rays sealed at their first lookup at or above it ---------------------
def _sun_world(seed):
    """Terrain that fights the sun clearance: stairs rising towards the sun (+x, +z), plateaus, thin slabs and pillars floating
    a few blocks above the surface along the sun direction of the blocks under them."""
    def fill(bm):
        rng = np.random.default_rng(seed)
        base = 8
        for x in range(16, 112):
            for z in range(16, 112):
                h = base + ((x + z) // (3 + seed % 3)) % 7 + (2 if (x // 9 + z // 11) % 3 == 0 else 0)
                for y in range(max(h - 2, 0), h):
                    bm.set(x, y, z, WATER)
        for _ in range(60):   # obstacles placed where SUN_DIR rays of nearby ground pass: (dx, dy, dz) ~ k * (0.75, 0.66, 0.75)
            x, z = int(rng.integers(20, 100)), int(rng.integers(20, 100))
            k = int(rng.integers(2, 14))
            ox, oz, oy = x + int(round(0.75 * k)) + int(rng.integers(-1, 2)), z + int(round(0.75 * k)) + int(rng.integers(-1, 2)), base + 8 + int(round(0.66 * k)) + int(rng.integers(-2, 3))
            for a in range(int(rng.integers(1, 4))):
                for b in range(int(rng.integers(1, 4))):
                    if ox + a < 126 and oz + b < 126:
                        bm.set(ox + a, oy, oz + b, WATER)
    return fill


@pytest.mark.parametrize("seed", range(4))
def test_sun_clearance_against_adversarial_terrain_and_step_caps(uvt, oracle, scene_factory, seed):
    down = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float32)
    with uvt.Context(0, map_dim=128, hit_buffer=True) as ctx:
        sc = scene_factory(128, _sun_world(seed), ctx=ctx)
        W, H = 224, 160
        ctx.resize(W, H)
        cams = [oracle.make_camera((64.0, 60.0, 64.0), down, 1.6), oracle.make_camera((30.0, 40.0, 30.0), pitch_yaw_matrix(uvt, 0.7, np.pi / 4)),
                oracle.make_camera((100.0, 35.0, 90.0), pitch_yaw_matrix(uvt, 0.5, 3.9))]
        for shadow_cap in (48, 17, 96, 48):   # growing the cap rebuilds the map (ensure_sun); shrinking keeps the larger one
            ctx.set_max_steps(192, shadow_cap)
            prm = oracle.params(128, shadow_max_steps=shadow_cap)
            for cam in cams:
                g = gpu_render(ctx, cam)
                r = oracle.render(sc.oracle_world, cam, W, H, prm)
                assert np.array_equal(g["illumination"], r["illumination"]), (seed, shadow_cap)
                assert len(np.unique(r["illumination"])) == 3 or (r["illumination"] != 0).any()
                assert ctx.count_pass("secondary") == {**r["secondary_counters"]}
        # a new obstacle far along the sun direction of lit ground, published incrementally: the map follows (update_sun)
        for x, y, z in ((70, 30, 72), (71, 30, 72), (70, 31, 73), (90, 25, 88)):
            sc.bm.set(x, y, z, WATER)
        sc.bm.bind(9)
        inc = ctx.world_layout_checksum()
        world = oracle.World(128, sc.bm.chunks().copy(), sc.bm.bricks().copy(), oracle.atlas_from_models(sc.models))
        for cam in cams:
            assert np.array_equal(gpu_render(ctx, cam)["illumination"], oracle.render(world, cam, W, H)["illumination"])
        sc.bm.mark_dirty()
        sc.bm.bind(9)   # full commit of the same world
        assert ctx.world_layout_checksum() == inc


def test_frame_in_row_chunks_equals_whole_frame_launches(uvt, oracle, w1):
    """uvt_set_frame_chunks: the same kernels over row chunks on two streams — every G-buffer image, the hit buffer and the
    frame are the whole-frame launches' bit for bit (also under a band partition and with the entity passes on)."""
    ctx, sc = w1
    ctx.set_layout("compact")
    cam = camera_k1(uvt, oracle)
    try:
        for part in (None, (16, 4, 1)):
            if part:
                ctx.set_partition(*part)
            ctx.resize(640, 360)
            ctx.set_frame_chunks(1)
            a = gpu_render(ctx, cam, three_pass=False)
            for n in (2, 3, 8):
                ctx.set_frame_chunks(n)
                b = gpu_render(ctx, cam, three_pass=False)
                for k in ("albedo", "normal", "illumination", "frame"):
                    assert np.array_equal(a[k], b[k]), (part, n, k)
                assert np.array_equal(a["position"].view(np.uint32), b["position"].view(np.uint32))
                assert np.array_equal(a["hits"], b["hits"])
        ctx.set_partition(8, 1, 0)
        ctx.resize(320, 180)
        ctx.set_entity_mode("models")
        p, pitch, yaw = ENTITY_CAMERAS[0]
        ecam = oracle.make_camera(p, pitch_yaw_matrix(uvt, pitch, yaw))
        ctx.set_frame_chunks(2)
        r = oracle.render(sc.oracle_world, ecam, 320, 180, oracle.params(512, ent=oracle.entities("models")))
        assert_frame_parity(gpu_render(ctx, ecam, three_pass=False), r)
    finally:
        ctx.set_frame_chunks(1)
        ctx.set_entity_mode("boxes")
        ctx.set_partition(8, 1, 0)
