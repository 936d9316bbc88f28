"""Pins the CPU oracle (oracle/oracle.c): hand-derived known answers, an
independent pure-Python restatement written from the shader text (oracle/pyref.py), committed
golden outputs, and domain properties.  The reference ships no golden vectors for this path
 and cannot run here (§8c), so these are what the oracle is pinned to."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, camera_k0, camera_k1

WATER = 0x1000000D
Y_DOWN = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)  # v4 = (x, -1, y, 1) at uv = 0


# ---- A.7 known answers --------------------------------------------------------------------
def test_kat_empty_world_every_pixel_misses(oracle, scene_factory):
    sc = scene_factory(512, None, key="empty")
    r = oracle.render(sc.oracle_world, camera_k0(oracle), 32, 18)
    assert (r["hits"]["face"] == 0).all() and (r["hits"]["color"] == 0).all()
    assert (r["position"] == -1.0).all() and (r["normal"] == 0xFFFFFFFF).all()
    assert (r["illumination"] == 0).all()
    # camera at (256,25,256): a ray cannot reach a map face within 192 trips of <= 8 sub-voxels... except steep ones
    trips = r["hits"]["trips"]
    assert trips.max() == 192 and (r["hits"]["exit_kind"][trips == 192] == 1).all()
    assert r["primary_counters"]["t_chunk"] == 0 and r["primary_counters"]["t_block"] == 0
    assert r["secondary_counters"]["early_out"] == 32 * 18 and r["secondary_counters"]["rays"] == 0


def test_kat_single_water_block_hand_trace(oracle, scene_factory):
    """A.7(ii): trips 1-3 block steps, trip 4 enters the water block at y'=7 (empty), trip 5 y'=6, trip 6 hits."""
    sc = scene_factory(512, lambda bm: bm.set(256, 0, 256, WATER), key="onewater")
    cam = oracle.make_camera((256.5, 4.0, 256.5), Y_DOWN)
    o, d, s = oracle.primary_ray(cam, 64, 64, 32, 32, 512)
    assert d.tolist() == [0.0, np.float32(-1 / np.sqrt(2)), 0.0]           # 4-component normalise: |d| < 1
    assert np.allclose(s, [256.499, 3.999, 256.499])
    h = oracle.trace_map(sc.oracle_world, s, d, 192)
    assert h["trips"] == 6 and h["p"] == (2052, 5, 2052) and h["data"] == 0xFFFFCC99
    assert h["face"] == 4 and h["normal"] == (0.0, 1.0, 0.0) and h["block"] == WATER
    assert np.allclose(h["hit_pos"], (2052.0286, 5.999, 2052.0286), atol=1e-4)
    assert (h["t_in"], h["t_chunk"], h["t_block"]) == (6, 6, 3)
    r = oracle.primary(sc.oracle_world, cam, 64, 64)
    assert r["position"][32, 32].tolist() == [256.625, 0.75, 256.625, 1.0]   # ceil(hit_pos)/8
    assert r["normal"][32, 32] == 0xFF00FF00 and r["albedo"][32, 32] == 0xFFFFCC99


def test_kat_illumination_bytes(oracle, scene_factory):
    """A.7(iii),(iv): illum = (192,168,192,77) lit / (192,168,192,0) shadowed; 0.3*255 = 76.5000030 -> 77."""
    assert np.float32(0.3) * np.float32(255) == np.float32(76.5)              # exactly .5 in fp32 ...
    assert int(np.float32(0.3) * np.float32(255) + np.float32(0.5)) == 77       # ... the +0.5/truncate UNORM rule gives 77

    def fill(bm):
        for x in range(250, 262):
            for z in range(250, 262):
                bm.set(x, 0, z, WATER)
        bm.set(258, 2, 258, WATER)  # an occluder toward the sun from around (256,0,256)
    sc = scene_factory(512, fill, key="slab+occluder")
    cam = oracle.make_camera((256.5, 4.0, 256.5), Y_DOWN)
    r = oracle.render(sc.oracle_world, cam, 64, 64)
    vals = set(np.unique(r["illumination"]).tolist())
    lit, shadowed = 192 | 168 << 8 | 192 << 16 | 77 << 24, 192 | 168 << 8 | 192 << 16
    assert vals <= {0, lit, shadowed} and lit in vals and shadowed in vals


def test_kat_ray_starting_inside_solid_reports_x_face(oracle, scene_factory):
    """A.5: initial minIdx = 0 -> a ray that starts inside a solid sub-voxel reports an X face."""
    sc = scene_factory(512, lambda bm: bm.set(256, 0, 256, WATER), key="onewater")
    h = oracle.trace_map(sc.oracle_world, (256.5, 0.3, 256.5), (0.0, -0.7, 0.0), 192)
    assert h["trips"] == 1 and h["face"] == 1  # d.x == 0 is patched to +0.001 -> positivity 1 -> faceId 2-1 = 1


def test_trace_leaves_map(oracle, scene_factory):
    sc = scene_factory(512, None, key="empty")
    h = oracle.trace_map(sc.oracle_world, (511.5, 3.0, 100.0), (0.9, 0.0, 0.0), 192)
    assert h["exit_kind"] == 2 and h["data"] == 0 and h["trips"] < 5 and h["hit_pos"] == (-1.0, -1.0, -1.0)


# ---- oracle.c vs the independent Python restatement -----------------------------------------
def _pyworld(oracle, sc, atlas):
    from oracle import pyref
    return pyref.World(sc.dim, sc.chunks, sc.bricks, atlas)


@pytest.mark.parametrize("cam_name", ["k0", "k1"])
def test_c_oracle_matches_python_restatement_primary(uvt, oracle, world512, atlas, cam_name):
    from oracle import pyref
    cam = camera_k0(oracle) if cam_name == "k0" else camera_k1(uvt, oracle)
    W, H = 48, 27
    pw = _pyworld(oracle, world512, atlas)
    r = oracle.primary(world512.oracle_world, cam, W, H)
    rng = np.random.default_rng(7)
    n_hit = 0
    for px, py in zip(rng.integers(0, W, 70), rng.integers(0, H, 70)):
        o, d, s = pyref.primary_ray(cam["cam_pos"], cam["cam_mat"], cam["fov"], W, H, int(px), int(py), 512)
        oc, dc, sc_ = oracle.primary_ray(cam, W, H, int(px), int(py), 512)
        assert np.array_equal(d, dc) and np.array_equal(s, sc_)
        h = pyref.trace_map(pw, s, d, 192)
        rec = r["hits"][py, px]
        assert (h["p"][0] & 0xFFFFFFFF, h["p"][1] & 0xFFFFFFFF, h["p"][2] & 0xFFFFFFFF) == (rec["px"], rec["py"], rec["pz"])
        assert (h["face"], h["block"], h["data"], h["trips"], h["exit_kind"]) == (rec["face"], rec["block"], rec["color"], rec["trips"], rec["exit_kind"])
        n_hit += h["data"] != 0
    assert n_hit > 10


def test_c_oracle_matches_python_restatement_shadow(uvt, oracle, world512, atlas):
    from oracle import pyref
    cam = camera_k1(uvt, oracle)
    W, H = 48, 27
    pw = _pyworld(oracle, world512, atlas)
    r = oracle.render(world512.oracle_world, cam, W, H)
    ys, xs = np.nonzero(r["hits"]["face"])
    rng = np.random.default_rng(3)
    pick = rng.choice(len(ys), 60, replace=False)
    seen = set()
    for i in pick:
        y, x = ys[i], xs[i]
        o = pyref.shadow_origin(r["position"][y, x, :3], int(r["normal"][y, x]))
        h = pyref.trace_map(pw, o, pyref.SUN_DIR, 48)
        dist = np.sqrt(np.float32(sum(np.float32(a) for a in ((o - np.array(h["hit_pos"], np.float32) / np.float32(8)) ** 2))))
        ent = pyref.trace_entities(o, pyref.SUN_DIR, np.float32(dist))
        expect_a = 0 if (ent or h["data"] != 0) else 77
        assert int(r["illumination"][y, x]) >> 24 == expect_a
        seen.add(expect_a)
    assert seen == {0, 77}


def test_random_rays_c_vs_python(oracle, world64, atlas):
    """Arbitrary origins/directions in a small world, including axis-aligned and zero components."""
    from oracle import pyref
    pw = pyref.World(64, world64.chunks, world64.bricks, atlas)
    rng = np.random.default_rng(11)
    for i in range(120):
        o = rng.uniform(-2, 66, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if i % 7 == 0:
            d[rng.integers(0, 3)] = 0.0
        if i % 11 == 0:
            d = np.array([0, -1, 0], np.float32)
        a = oracle.trace_map(world64.oracle_world, o, d, 64)
        b = pyref.trace_map(pw, o, d, 64)
        for k in ("data", "face", "block", "trips", "exit_kind", "t_in", "t_chunk", "t_block"):
            assert a[k] == b[k], (i, k, a, b)
        assert a["hit_pos"] == b["hit_pos"]
        assert tuple(x & 0xFFFFFFFF for x in b["p"]) == a["p"]


# ---- entities, sky, blit ----------------------------------------------------------------------
def test_trace_entities_quirks(oracle):
    """A.5: boxes intersect as LINES (no t >= 0 test), selection gated by distance < maxDistance."""
    L = oracle.lib()

    def f(o, d, m):
        o, d = np.asarray(o, np.float32), np.asarray(d, np.float32)  # keep the arrays alive across the call
        return L.orc_trace_entities(o.ctypes.data, d.ctypes.data, float(m))

    assert f((256.5, 10.0, 256.5), (0.001, 1.0, 0.001), 1000.0) == 1      # box above along +y
    assert f((256.5, 30.0, 256.5), (0.001, 1.0, 0.001), 1000.0) == 1      # box BEHIND the origin still counts
    assert f((256.5, 10.0, 256.5), (0.001, 1.0, 0.001), 5.0) == 0         # too far: distance >= maxDistance
    assert f((100.5, 10.0, 100.5), (0.001, 1.0, 0.001), 1000.0) == 0      # line misses every box


def test_sky_dome(oracle):
    sun = np.array([7.52185881e-01, 6.58950984e-01, 7.52185881e-01], np.float32)
    c = oracle.sky_dome2(sun)  # looking at the sun: sun = 1
    k = sun[1] * np.float32(0.2)
    assert np.allclose(c, [0.6 - k + 0.075 + 0.4 + 0.2, 0.71 - k * 0.5 + 0.075 + 0.24 + 0.08, 0.75 - k + 0.075 + 0.04 + 0.04], atol=1e-5)
    c = oracle.sky_dome2(-sun)  # away: sun clamps to 0
    assert np.allclose(c, [0.6 + k + 0.075, 0.71 + k * 0.5 + 0.075, 0.75 + k + 0.075], atol=1e-6)


def test_blit_formula(oracle):
    W, H = 64, 36
    albedo = np.full((H, W), 0xFF808080, np.uint32)
    normal = np.full((H, W), 0xFF00FF00, np.uint32)
    position = np.ones((H, W, 4), np.float32)
    illum = np.full((H, W), 192 | 168 << 8 | 192 << 16 | 77 << 24, np.uint32)
    illum[:, : W // 2] &= 0x00FFFFFF  # shadowed half: alpha 0 -> colour = albedo
    f = oracle.blit(albedo, normal, position, illum)
    x, y = 10, 20
    tx, ty = (x + 0.5) / W, (y + 0.5) / H
    grad = (tx * (1 - tx) * ty * (1 - ty) * 15.0) ** 0.18
    assert abs((int(f[y, x]) & 255) - round(grad * (128 / 255) * 255)) <= 1
    lit = int(f[y, W - 1 - x]) & 255
    assert lit > (int(f[y, x]) & 255)       # same vignette by symmetry, brighter where lit
    # crosshair: |texPos - 0.5| <= 0.002 never holds at 64x36 (nearest pixel centre is 0.0078 away)
    W2, H2 = 640, 360
    f2 = oracle.blit(np.zeros((H2, W2), np.uint32), np.zeros((H2, W2), np.uint32), np.zeros((H2, W2, 4), np.float32), np.zeros((H2, W2), np.uint32))
    ys, xs = np.nonzero(f2 & 0xFFFFFF)
    assert len(ys) > 0 and np.all(np.hypot((xs + 0.5) / W2 - 0.5, (ys + 0.5) / H2 - 0.5) <= 0.002 + 1e-6)


# ---- properties ---------------------------------------------------------------------------------
def test_counters_consistent_with_hit_buffer(uvt, oracle, world512):
    r = oracle.primary(world512.oracle_world, camera_k1(uvt, oracle), 160, 90)
    c, hits = r["counters"], r["hits"]
    assert c["rays"] == 160 * 90 and c["hits"] == int((hits["face"] != 0).sum())
    assert c["t_in"] == int(hits["trips"].astype(np.int64).sum())   # every counted trip passed the bounds test
    assert c["t_block"] <= c["t_chunk"] <= c["t_in"]
    assert oracle.algorithmic_bytes(c, 160 * 90, "primary") == 4 * (c["t_in"] + c["t_chunk"] + c["t_block"]) + 24 * 160 * 90


def test_thread_count_does_not_change_results(uvt, oracle, world512):
    cam = camera_k1(uvt, oracle)
    n = oracle.num_threads()
    a = oracle.render(world512.oracle_world, cam, 96, 54)
    oracle.set_num_threads(1)
    b = oracle.render(world512.oracle_world, cam, 96, 54)
    oracle.set_num_threads(n)
    for k in ("albedo", "normal", "position", "illumination", "frame"):
        assert np.array_equal(a[k], b[k])
    assert a["primary_counters"] == b["primary_counters"] and a["secondary_counters"] == b["secondary_counters"]


def test_hit_geometry_invariants(uvt, oracle, world512):
    r = oracle.primary(world512.oracle_world, camera_k1(uvt, oracle), 160, 90)
    h = r["hits"]
    hit = h["face"] != 0
    assert hit.mean() > 0.3
    assert ((h["px"][hit] < 4096) & (h["py"][hit] < 4096) & (h["pz"][hit] < 4096)).all()
    assert (h["color"][hit] >> 24 == 0xFF).all() and (h["block"][hit] != 0).all()
    assert (h["trips"][hit] >= 1).all() and (h["trips"] <= 192).all()
    assert (h["exit_kind"][hit] == 0).all() and (h["exit_kind"][~hit] >= 1).all()
    # stored position = ceil(hit_pos)/8 lies within one sub-voxel of the hit cell
    pos = r["position"][hit][:, :3] * 8
    cell = np.stack([h["px"][hit], h["py"][hit], h["pz"][hit]], -1).astype(np.float32)
    assert (np.abs(pos - cell) <= 1.0).all()
    assert (h["distance"][hit] > 0).all() and (h["distance"][~hit] == -1).all()


# ---- committed golden outputs -------------------------------------------------------------------
@pytest.mark.parametrize("name", ["k0_96x54", "k1_96x54"])
def test_oracle_reproduces_committed_golden(uvt, oracle, world512, name):
    g = np.load(os.path.join(GOLDEN, f"oracle_w1_{name}.npz"))
    cam = camera_k0(oracle) if name.startswith("k0") else camera_k1(uvt, oracle)
    r = oracle.render(world512.oracle_world, cam, 96, 54)
    for k in ("albedo", "normal", "position", "illumination", "frame"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["hits"].view(np.uint8), g["hits"].view(np.uint8).reshape(r["hits"].view(np.uint8).shape))
    assert [r["primary_counters"][k] for k in ("rays", "t_in", "t_chunk", "t_block", "hits")] == g["primary_counters"].tolist()


def test_pixel_list_api_equals_the_full_frame(oracle, world512, uvt):
    """orc_primary_pixels (sampled parity of the 4K / 8K frames) is the same code path as the full-frame pass."""
    cam = camera_k1(uvt, oracle)
    W, H = 200, 120
    full = oracle.render(world512.oracle_world, cam, W, H)
    rng = np.random.default_rng(3)
    xs, ys = rng.integers(0, W, 3000), rng.integers(0, H, 3000)
    p = oracle.primary_pixels(world512.oracle_world, cam, W, H, xs, ys)
    assert np.array_equal(p["hits"], full["hits"][ys, xs])
    for k in ("albedo", "normal"):
        assert np.array_equal(p[k], full[k][ys, xs])
    assert np.array_equal(p["position"].view(np.uint32), full["position"][ys, xs].view(np.uint32))
    s = oracle.secondary(world512.oracle_world, p["normal"][None, :], p["position"][None, :, :])
    assert np.array_equal(s["illumination"][0], full["illumination"][ys, xs])


# ---- entities done properly: traceEntities with map.glsl:203-248 live ---------------------------
def _entity_rays(rng, n, positions, edge):
    """Rays aimed at (or near) the entity boxes from around them; every 7th has a zero direction component (the text
    does not patch zeros here: sign(0) = 0, infinite reciprocal)."""
    P = np.asarray(positions, np.float32)
    lo, hi = P.min(0) - 6, P.max(0) + edge + 6
    for i in range(n):
        o = rng.uniform(lo, hi).astype(np.float32)
        tgt = P[rng.integers(len(P))] + rng.uniform(-0.2, edge + 0.2, 3).astype(np.float32)
        d = tgt - o
        d = (d / np.linalg.norm(d)).astype(np.float32)
        if i % 7 == 0:
            d[rng.integers(3)] = 0.0
        yield o, d


def test_kat_entity_model_axis_ray(oracle, scene_factory, atlas):
    """Hand trace of map.glsl:203-230 for a full 8^3 model: the ray (1,0,0) from (250, 21.5, 256.5) meets entity 3 at
    (251,21,256) with tNear = 1: gridsCoords = ivec3((-0.001, 0.499, 0.499) * 8) = (0, 3, 3), withinGridCoords = (0, 1, 1),
    the first lookup is voxel (0, 4, 4), minIdx is still 0 so faceId = 2 - rayPositivity.x = 1, hit_pos = pos + (0, 4, 4) / 8."""
    sc = scene_factory(512, None, key="empty")
    full = np.full(512, 0xFF112233, np.uint32)
    ent = oracle.entities("models", model=full, size=8)
    h = oracle.trace_entities_ex(sc.oracle_world, ent, (250.0, 21.5, 256.5), (1.0, 0.0, 0.0), 100.0)
    assert h["data"] == 0xFF112233 and h["entity"] == 3 and h["trips"] == 1 and h["p"] == (0, 4, 4)
    assert h["face"] == 1 and h["normal"] == (-1.0, 0.0, 0.0) and h["hit_pos"] == (251.0, 21.5, 256.5)
    # boxes mode (the reference as it runs): data = 0xFFFFFFFF, hit_pos = positions[id]  (map.glsl:199)
    b = oracle.trace_entities_ex(sc.oracle_world, oracle.entities("boxes"), (250.0, 21.5, 256.5), (1.0, 0.0, 0.0), 100.0)
    assert b["data"] == 0xFFFFFFFF and b["hit_pos"] == (251.0, 21.0, 256.0)
    # maxDistance culls on distance(rayOrigin, positions[i]) (the box CORNER), strictly: 1.118... away
    assert oracle.trace_entities_ex(sc.oracle_world, ent, (250.0, 21.5, 256.5), (1.0, 0.0, 0.0), 1.0)["data"] == 0
    # an empty model and the quirk of the text: rayDir is NOT zero-patched here, so t.y = (0 - 1) * (1 / 0) = -inf wins the
    # argmin, raySign.y = 0 moves nothing, `within` turns NaN and the ray idles inside the grid until the 64-trip cap
    e0 = oracle.entities("models", model=np.zeros(512, np.uint32), size=8)
    h0 = oracle.trace_entities_ex(sc.oracle_world, e0, (250.0, 21.5, 256.5), (1.0, 0.0, 0.0), 100.0)
    assert h0["data"] == 0 and h0["trips"] == 64
    # without zero components it crosses the 8 voxels of the box and leaves: 8 x-steps plus the y/z steps on the way
    h1 = oracle.trace_entities_ex(sc.oracle_world, e0, (250.0, 21.5, 256.5), (1.0, 0.01, 0.02), 100.0)
    assert h1["data"] == 0 and h1["exit_kind"] == 2 and 8 <= h1["trips"] <= 10


@pytest.mark.parametrize("case", ["atlas8", "chicken32", "custom16"])
def test_entity_models_c_vs_python_restatement(oracle, scene_factory, models, case):
    from oracle import pyref
    sc = scene_factory(512, None, key="empty")
    rng = np.random.default_rng({"atlas8": 1, "chicken32": 2, "custom16": 3}[case])
    if case == "atlas8":      # the text as written: texels [0,8)^3 of the atlas = the first block model
        ent, model, size, steps, pos = oracle.entities("models"), models[0].reshape(8, 8, 8), 8, 64, oracle.ENTITY_POSITIONS
    elif case == "chicken32":
        m = np.load(os.path.join(GOLDEN, "chicken_32.npy"))
        pos = [(100.0, 30.0, 100.0), (110.5, 31.25, 97.0)]
        ent, model, size, steps = oracle.entities("models", positions=pos, model=m, size=32, max_steps=256), m.reshape(32, 32, 32), 32, 256
    else:
        m = (rng.random(16 ** 3) < 0.04).astype(np.uint32) * rng.integers(1, 2 ** 32, 16 ** 3, dtype=np.uint32)
        pos = [(20.0, 5.0, 20.0), (21.0, 5.5, 22.5), (23.0, 4.0, 20.0)]
        ent, model, size, steps = oracle.entities("models", positions=pos, model=m, size=16, max_steps=40), m.reshape(16, 16, 16), 16, 40
    hits = 0
    for o, d in _entity_rays(rng, 1500, pos, size / 8):
        a = oracle.trace_entities_ex(sc.oracle_world, ent, o, d, 60.0)
        b = pyref.trace_entities_models(model, o, d, np.float32(60.0), positions=pos, size=size, max_steps=steps)
        assert a["data"] == b["data"] and a["trips"] == b["trips"], (o, d, a, b)
        if a["data"]:
            hits += 1
            assert (a["face"], a["p"], a["entity"]) == (b["face"], b["p"], b["entity"])
            assert np.array_equal(np.array(a["hit_pos"], np.float32), np.array(b["hit_pos"], np.float32))
    assert hits > 300, hits


def test_entity_composite_frame_properties(uvt, oracle, world512):
    """primary.comp.glsl:45-54 made live: entity pixels carry the model's colour, a world-space position inside the
    entity's box and exit_kind 3; every other pixel is the frame without entities; the shadow pass then treats entity
    surfaces like terrain.  The boxes mode reproduces the default params bit for bit."""
    cam = oracle.make_camera((258.0, 25.0, 262.0), np.asarray(_pitch_yaw(uvt, 0.5, 3.6)))
    W, H = 160, 90
    base = oracle.render(world512.oracle_world, cam, W, H)
    boxes = oracle.render(world512.oracle_world, cam, W, H, oracle.params(512, ent=oracle.entities("boxes")))
    for k in ("albedo", "normal", "illumination", "frame"):
        assert np.array_equal(base[k], boxes[k]), k
    r = oracle.render(world512.oracle_world, cam, W, H, oracle.params(512, ent=oracle.entities("models")))
    ent_px = r["hits"]["exit_kind"] == 3
    assert 200 < ent_px.sum() < W * H // 2, ent_px.sum()
    for k in ("albedo", "normal"):
        assert np.array_equal(r[k][~ent_px], base[k][~ent_px])
    p = r["position"][ent_px][:, :3]
    ids = r["hits"]["block"][ent_px] & 0xFF
    corner = np.asarray(oracle.ENTITY_POSITIONS, np.float32)[ids]
    assert ((p >= corner - 1e-3) & (p <= corner + 1 + 1e-3)).all()
    assert (r["hits"]["color"][ent_px] != 0).all() and set(np.unique(r["hits"]["face"][ent_px])) <= {1, 2, 3, 4, 5, 6}
    assert (r["illumination"][ent_px] != 0).all()  # entity surfaces shoot shadow rays


def _pitch_yaw(uvt, pitch, yaw):
    from conftest import pitch_yaw_matrix
    return pitch_yaw_matrix(uvt, pitch, yaw)
