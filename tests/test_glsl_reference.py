"""The CPU oracle against the reference's OWN shader text.

oracle/_ref/libglslref.so is assets/shaders/{primary.comp,secondary.comp,blit.fragment}.glsl (+ camera/map/rng.glsl) of
the reference checkout, translated by syntactic rewrites only (oracle/glsl_ref/translate.py) and compiled for the CPU
against a GLSL-in-C++ shim (oracle/glsl_ref/glsl_shim.h).  These tests hold oracle/oracle.c — the checker of every GPU
parity test — to that library bit for bit: every G-buffer image, the illumination image and the final frame, and
traceMap / traceEntities ray by ray.  They run wherever the library exists (built here from /root/reference; the
prebuilt files travel to the GPU box)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, camera_k0, camera_k1, pitch_yaw_matrix

WATER = 0x1000000D


@pytest.fixture(scope="module")
def glslref():
    from oracle import glslref as g
    if not g.available():
        pytest.skip("oracle/_ref is not built and the reference checkout is absent")
    return g


def assert_same_frame(r, g):
    for k in ("albedo", "normal", "illumination", "frame"):
        assert np.array_equal(r[k], g[k]), f"{k}: {(r[k] != g[k]).sum()} texels differ from the reference shader text"
    assert np.array_equal(r["position"].view(np.uint32), g["position"].view(np.uint32))


POSES = [((256.0, 25.0, 256.0), None, np.pi / 2), ((200.0, 48.0, 140.0), (0.35, 0.6), np.pi / 2), ((262.0, 30.0, 262.0), (0.6, 5 * np.pi / 4), np.pi / 2),
         ((249.0, 27.0, 262.0), (0.9, 2.2), 1.1), ((100.5, 60.0, 400.25), (-0.4, 4.0), 2.0), ((-20.0, 40.0, 256.0), (0.2, np.pi / 2), np.pi / 2),
         ((256.0, 600.0, 256.0), (1.2, 0.3), np.pi / 2), ((5.0, 18.0, 5.0), (0.1, 3.9), 1.5), ((256.0, 30.0, 256.0), "down", np.pi / 2)]


def _cam(uvt, oracle, pose):
    p, rot, fov = pose
    if rot == "down":
        return oracle.make_camera(p, np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float32), fov)
    return oracle.make_camera(p, None if rot is None else pitch_yaw_matrix(uvt, *rot), fov)


@pytest.mark.parametrize("i", range(len(POSES)))
def test_oracle_frames_equal_the_reference_shader_text_w1(uvt, oracle, glslref, world512, i):
    cam = _cam(uvt, oracle, POSES[i])
    W, H = (213, 120) if i % 2 else (256, 144)   # 213 x 120: ragged against the 32 x 32 work groups
    r = oracle.render(world512.oracle_world, cam, W, H)
    assert_same_frame(r, glslref.render(world512.oracle_world, cam, W, H))


def test_oracle_frames_equal_the_reference_shader_text_720p(uvt, oracle, glslref, world512):
    """BASELINE config 1 at full size (1280x720, primary + shadow rays, camera K0) and K1."""
    for cam in (camera_k0(oracle), camera_k1(uvt, oracle)):
        r = oracle.render(world512.oracle_world, cam, 1280, 720)
        assert_same_frame(r, glslref.render(world512.oracle_world, cam, 1280, 720))
        assert len(np.unique(r["illumination"])) == 3


def test_small_and_hand_made_worlds(uvt, oracle, glslref, scene_factory):
    cases = [(512, None, "empty"), (512, lambda bm: bm.set(256, 0, 256, WATER), "onewater"), (64, "procgen", None), (128, "procgen", None)]
    for dim, fill, key in cases:
        sc = scene_factory(dim, fill, key=key)
        c = dim / 2
        for cam in (oracle.make_camera((c + 0.5, min(dim - 2, 30.0), c + 0.5), pitch_yaw_matrix(uvt, 0.7, 0.4)),
                    oracle.make_camera((256.5, 4.0, 256.5) if dim == 512 else (c, 12.0, c), np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float32))):
            r = oracle.render(sc.oracle_world, cam, 96, 54, oracle.params(dim))
            assert_same_frame(r, glslref.render(sc.oracle_world, cam, 96, 54))


def test_scaled_world_w4(uvt, oracle, glslref, scene_factory):
    """MAP_DIMENSION = 2048 (translate.py rule R7): the world of BASELINE configs 3-5, camera K1 of the bench and a pose from the sweep."""
    sc = scene_factory(2048, "procgen")
    cams = [uvt.scenes.camera_k1(2048), uvt.scenes.sweep_poses(2048, 3)[2]]
    for cam in cams:
        r = oracle.render(sc.oracle_world, cam, 240, 135, oracle.params(2048))
        assert_same_frame(r, glslref.render(sc.oracle_world, cam, 240, 135))
        assert (r["hits"]["face"] != 0).mean() > 0.2


def test_trace_map_ray_by_ray(oracle, glslref, world64, world512):
    rng = np.random.default_rng(11)
    n_hit = 0
    for sc, dim in ((world64, 64), (world512, 512)):
        for i in range(1500):
            o = rng.uniform(-4, dim + 4, 3).astype(np.float32)
            if i % 3:
                o[1] = np.float32(rng.uniform(0, 40))
            d = rng.normal(size=3).astype(np.float32)
            d /= np.float32(np.linalg.norm(d))
            if i % 11 == 0:
                d[rng.integers(3)] = 0.0      # the zero patch of map.glsl:85-90
            steps = int(rng.choice([1, 7, 48, 64, 192]))
            a = oracle.trace_map(sc.oracle_world, o, d, steps)
            b = glslref.trace_map(sc.oracle_world, o, d, steps)
            assert a["data"] == b["data"], (o, d, steps)
            assert np.array_equal(np.array(a["hit_pos"], np.float32), np.array(b["hit_pos"], np.float32))
            assert a["normal"] == b["normal"]
            n_hit += a["data"] != 0
    assert n_hit > 300


def test_trace_entities_live_behaviour(oracle, glslref, world512):
    """map.glsl:172-201 as it runs: data = 0xFFFFFFFF and hit_pos = positions[id] when the LINE meets the best box."""
    rng = np.random.default_rng(5)
    ent = oracle.entities("boxes")
    n = 0
    for i in range(3000):
        o = np.array([254 + rng.uniform(-8, 10), 21.5 + rng.uniform(-4, 8), 258 + rng.uniform(-8, 10)], np.float32)
        tgt = np.array(oracle.ENTITY_POSITIONS[rng.integers(5)], np.float32) + rng.uniform(-0.5, 1.5, 3).astype(np.float32)
        d = tgt - o
        d = (d / np.linalg.norm(d)).astype(np.float32)
        if i % 9 == 0:
            d[rng.integers(3)] = 0.0
        maxd = np.float32(rng.choice([0.5, 3.0, 8.0, 100.0]))
        a = oracle.trace_entities_ex(world512.oracle_world, ent, o, d, maxd)
        b = glslref.trace_entities(world512.oracle_world, o, d, maxd)
        assert a["data"] == b["data"], (o, d, maxd)
        if a["data"]:
            n += 1
            assert a["hit_pos"] == b["hit_pos"]
        assert bool(a["data"]) == bool(oracle.lib().orc_trace_entities(o.ctypes.data, d.ctypes.data, float(maxd)))
    assert n > 500


def test_entity_models_against_the_dead_code_made_live(uvt, oracle, glslref, world512):
    """the library built from the text with map.glsl:199 deleted and primary.comp.glsl:47-54 uncommented
    (translate.py rules E1, E2) against oracle.c's entity-model mode: rays and whole frames."""
    from oracle import glslref as g
    if not g.available("entities"):
        pytest.skip("entities variant not built")
    rng = np.random.default_rng(6)
    ent = oracle.entities("models")
    n = 0
    for i in range(2500):
        o = np.array([254 + rng.uniform(-6, 8), 21.5 + rng.uniform(-3, 6), 258 + rng.uniform(-6, 8)], np.float32)
        tgt = np.array(oracle.ENTITY_POSITIONS[rng.integers(5)], np.float32) + rng.uniform(0, 1, 3).astype(np.float32)
        d = tgt - o
        d = (d / np.linalg.norm(d)).astype(np.float32)
        if i % 7 == 0:
            d[rng.integers(3)] = 0.0
        a = oracle.trace_entities_ex(world512.oracle_world, ent, o, d, 50.0)
        b = g.trace_entities(world512.oracle_world, o, d, 50.0, variant="entities")
        assert a["data"] == b["data"], (o, d)
        if a["data"]:
            n += 1
            assert np.array_equal(np.array(a["hit_pos"], np.float32), np.array(b["hit_pos"], np.float32)) and a["normal"] == b["normal"]
    assert n > 1500
    prm = oracle.params(512, ent=ent)
    for p, pitch, yaw in (((258.0, 25.0, 262.0), 0.5, 3.6), ((262.0, 30.0, 262.0), 0.6, 5 * np.pi / 4), ((256.5, 21.5, 256.5), 0.2, 1.0)):
        cam = oracle.make_camera(p, pitch_yaw_matrix(uvt, pitch, yaw))
        r = oracle.render(world512.oracle_world, cam, 200, 112, prm)
        assert (r["hits"]["exit_kind"] == 3).sum() > 100
        assert_same_frame(r, g.render(world512.oracle_world, cam, 200, 112, variant="entities"))


def test_sky_dome(oracle, glslref):
    rng = np.random.default_rng(2)
    for _ in range(500):
        rd = rng.normal(size=3).astype(np.float32)
        rd /= np.float32(np.linalg.norm(rd)) * np.float32(rng.uniform(1.0, 1.3))
        assert np.array_equal(oracle.sky_dome2(rd), glslref.sky_dome2(rd)[:3])


@pytest.mark.parametrize("name", ["k0_96x54", "k1_96x54"])
def test_committed_golden_is_the_reference_shader_output(uvt, oracle, glslref, world512, name):
    """tests/golden/glslref_w1_*.npz (tools/make_golden.py) hold outputs of the reference's shader text; the GPU suite
    checks the CUDA path against them on the GPU box.  Here: the files are reproducible and equal the oracle's."""
    gold = np.load(os.path.join(GOLDEN, f"glslref_w1_{name}.npz"))
    cam = np.frombuffer(gold["camera"].tobytes(), dtype=oracle.CAMERA_DTYPE)[0]
    g = glslref.render(world512.oracle_world, cam, 96, 54)
    r = oracle.render(world512.oracle_world, cam, 96, 54)
    for k in ("albedo", "normal", "illumination", "frame"):
        assert np.array_equal(g[k], gold[k]) and np.array_equal(r[k], gold[k]), k
    assert np.array_equal(g["position"].view(np.uint32), gold["position"].view(np.uint32))


@pytest.mark.parametrize("seed", range(12))
def test_random_worlds_and_cameras(uvt, oracle, glslref, models, atlas, seed):
    """Seeded random 64^3 worlds (random blocks of every loaded model, solid and not, unloaded model ids, blocks on the map
    faces), random cameras inside / outside the map, random fov and frame size: oracle.c == the reference's shader text,
    every image, bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    dim = 64
    bm = uvt.voxel.VoxelBrickmap.init(dim)
    n_models = len(models)
    for _ in range(int(rng.integers(50, 1500))):
        x, y, z = (int(v) for v in rng.integers(0, dim, 3))
        if rng.random() < 0.1:
            x = int(rng.choice([0, dim - 1]))
        ty = int(rng.integers(0, n_models + (3 if seed % 4 == 0 else 0)))   # ids past the atlas: unloaded models read as empty
        bm.set(x, y, z, ty | (int(rng.integers(0, 2)) << 28))
    if seed % 3 == 0:   # a floor, so that shadows and sub-voxel grazing happen
        for x in range(dim):
            for z in range(dim):
                bm.set(x, 3, z, int(rng.integers(0, 6)) | (1 << 28))
    world = oracle.World(dim, bm.chunks().copy(), bm.bricks().copy(), atlas)
    for _ in range(3):
        inside = rng.random() < 0.7
        pos = rng.uniform(2, dim - 2, 3) if inside else rng.uniform(-30, dim + 30, 3)
        cam = oracle.make_camera(tuple(float(v) for v in pos), pitch_yaw_matrix(uvt, float(rng.uniform(-1.4, 1.4)), float(rng.uniform(0, 2 * np.pi))),
                                 float(rng.uniform(0.5, 2.4)))
        W, H = int(rng.integers(8, 120)), int(rng.integers(8, 90))
        r = oracle.render(world, cam, W, H, oracle.params(dim))
        assert_same_frame(r, glslref.render(world, cam, W, H))
