"""CPU parity oracle — TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/oracle.c (the C restatement of the reference GLSL path; see the
header of that file for what pins it).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package; the product
package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

HIT_DTYPE = np.dtype([("px", "<u4"), ("py", "<u4"), ("pz", "<u4"), ("block", "<u4"), ("color", "<u4"),
                      ("distance", "<f4"), ("trips", "<u2"), ("face", "u1"), ("exit_kind", "u1")])
assert HIT_DTYPE.itemsize == 28

CAMERA_DTYPE = np.dtype([("cam_pos", "<f4", 4), ("cam_mat", "<f4", 16), ("fov", "<f4"), ("_pad", "<f4", 3)])
assert CAMERA_DTYPE.itemsize == 96


class _World(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_uint32), ("chunks", ctypes.c_void_p), ("bricks", ctypes.c_void_p), ("atlas", ctypes.c_void_p)]


MAX_ENTITIES = 32
ENTITY_POSITIONS = ((256., 21., 256.), (251., 21., 259.), (253., 21., 256.), (251., 21., 256.), (257., 21., 261.))  # map.glsl:173-179


class _Entities(ctypes.Structure):
    _fields_ = [("mode", ctypes.c_uint32), ("n", ctypes.c_uint32), ("pos", (ctypes.c_float * 3) * MAX_ENTITIES),
                ("model", ctypes.c_void_p), ("size", ctypes.c_uint32), ("max_steps", ctypes.c_uint32)]


class _Params(ctypes.Structure):
    _fields_ = [("map_dim", ctypes.c_uint32), ("primary_max_steps", ctypes.c_uint32), ("shadow_max_steps", ctypes.c_uint32),
                ("epsilon", ctypes.c_float), ("entities", ctypes.c_uint32), ("ent", ctypes.POINTER(_Entities))]


class _Hit(ctypes.Structure):
    _fields_ = [("data", ctypes.c_uint32), ("hit_pos", ctypes.c_float * 3), ("normal", ctypes.c_float * 3),
                ("p", ctypes.c_uint32 * 3), ("face", ctypes.c_uint32), ("block", ctypes.c_uint32), ("trips", ctypes.c_uint32),
                ("exit_kind", ctypes.c_uint32), ("t_in", ctypes.c_uint32), ("t_chunk", ctypes.c_uint32), ("t_block", ctypes.c_uint32),
                ("distance", ctypes.c_float)]


class _Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("rays", "t_in", "t_chunk", "t_block", "hits", "early_out")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force=False):
    """Compile oracle.c -> liboracle.so (gcc, -ffp-contract=off, OpenMP)."""
    src = os.path.join(_HERE, "oracle.c")
    hdr = os.path.join(_HERE, "oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, env=env, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_trace_map.argtypes = [ctypes.POINTER(_World), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_Hit)]
        L.orc_trace_entities.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]
        L.orc_trace_entities.restype = ctypes.c_int
        L.orc_trace_entities_ex.argtypes = [ctypes.POINTER(_World), ctypes.POINTER(_Entities), ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_float, ctypes.POINTER(_Hit)]
        L.orc_primary_ray.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_primary.argtypes = [ctypes.POINTER(_World), ctypes.c_void_p, ctypes.POINTER(_Params), ctypes.c_uint32, ctypes.c_uint32,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(_Counters)]
        L.orc_primary_pixels.argtypes = [ctypes.POINTER(_World), ctypes.c_void_p, ctypes.POINTER(_Params), ctypes.c_uint32, ctypes.c_uint32,
                                         ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_secondary.argtypes = [ctypes.POINTER(_World), ctypes.POINTER(_Params), ctypes.c_uint32, ctypes.c_uint32,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(_Counters)]
        L.orc_blit.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_sky_dome2.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def atlas_from_models(models):
    """Place [n,512] model texels (x + 8y + 64z) into the 256^3 RGBA8 atlas exactly as
    VoxelModelAtlas.load_single_block_model does (src/engine/voxel.zig:100-110)."""
    models = np.ascontiguousarray(models, dtype=np.uint32).reshape(-1, 512)
    atlas = np.zeros((256, 256, 256), dtype=np.uint32)  # [z][y][x]
    for idx, m in enumerate(models):
        bx, by, bz = idx % 32, (idx // 32) % 32, (idx // 1024) % 1024
        atlas[bz * 8:bz * 8 + 8, by * 8:by * 8 + 8, bx * 8:bx * 8 + 8] = m.reshape(8, 8, 8)
    return atlas


def make_camera(pos, mat=None, fov=np.pi / 2):
    cam = np.zeros((), dtype=CAMERA_DTYPE)
    cam["cam_pos"][:3] = pos
    cam["cam_mat"] = np.eye(4, dtype=np.float32).reshape(16) if mat is None else np.asarray(mat, dtype=np.float32).reshape(16)
    cam["fov"] = np.float32(fov)
    return cam


class World:
    """The world as the reference shaders see it: chunk table, brick pool, 256^3 atlas."""

    def __init__(self, dim, chunks, bricks, atlas):
        self.dim = int(dim)
        self.chunks = np.ascontiguousarray(chunks, dtype=np.uint32).reshape(-1)
        self.bricks = np.ascontiguousarray(bricks, dtype=np.uint32).reshape(-1)
        self.atlas = np.ascontiguousarray(atlas, dtype=np.uint32).reshape(-1)
        assert self.chunks.size == (self.dim // 8) ** 3
        assert self.atlas.size == 256 ** 3
        if self.bricks.size == 0:
            self.bricks = np.zeros(512, dtype=np.uint32)
        assert int(self.chunks.max(initial=0)) * 512 <= self.bricks.size
        self._c = _World(self.dim, self.chunks.ctypes.data, self.bricks.ctypes.data, self.atlas.ctypes.data)


def entities(mode="models", positions=None, model=None, size=8, max_steps=64):
    """The entity set of traceEntities (map.glsl:172-248).  mode "boxes" = the reference as it runs (early return at :199),
    "models" = the code behind that return made live + the primary composite (primary.comp.glsl:45-54).  Defaults are the
    literals of the text; model=None reads texels [0,8)^3 of the atlas like `imageLoad(model, ivec3(pos) & 7)`."""
    e = _Entities()
    e.mode = {"boxes": 0, "models": 1}[mode]
    pos = np.asarray(ENTITY_POSITIONS if positions is None else positions, dtype=np.float32).reshape(-1, 3)
    assert 0 < len(pos) <= MAX_ENTITIES
    e.n = len(pos)
    for i, q in enumerate(pos):
        for k in range(3):
            e.pos[i][k] = float(q[k])
    e.size, e.max_steps = int(size), int(max_steps)
    if model is not None:
        m = np.ascontiguousarray(model, dtype=np.uint32).reshape(-1)
        assert m.size == size ** 3 and size in (8, 16, 32)
        e._keep = m
        e.model = m.ctypes.data
    else:
        assert size == 8
    return e


def params(map_dim, primary_max_steps=192, shadow_max_steps=48, epsilon=0.001, entities=True, ent=None):
    p = _Params(int(map_dim), int(primary_max_steps), int(shadow_max_steps), float(np.float32(epsilon)), 1 if entities else 0,
                ctypes.pointer(ent) if ent is not None else None)
    p._keep = ent
    return p


def trace_entities_ex(world, ent, origin, direction, max_distance, epsilon=0.001):
    o = np.asarray(origin, dtype=np.float32)
    d = np.asarray(direction, dtype=np.float32)
    h = _Hit()
    lib().orc_trace_entities_ex(ctypes.byref(world._c), ctypes.byref(ent), float(np.float32(epsilon)), o.ctypes.data, d.ctypes.data,
                                float(np.float32(max_distance)), ctypes.byref(h))
    return {"data": h.data, "hit_pos": tuple(h.hit_pos), "normal": tuple(h.normal), "p": tuple(h.p), "face": h.face,
            "entity": h.block, "trips": h.trips, "exit_kind": h.exit_kind}


def trace_map(world, origin, direction, max_steps):
    o = np.asarray(origin, dtype=np.float32)
    d = np.asarray(direction, dtype=np.float32)
    h = _Hit()
    lib().orc_trace_map(ctypes.byref(world._c), o.ctypes.data, d.ctypes.data, int(max_steps), ctypes.byref(h))
    return {"data": h.data, "hit_pos": tuple(h.hit_pos), "normal": tuple(h.normal), "p": tuple(h.p), "face": h.face,
            "block": h.block, "trips": h.trips, "exit_kind": h.exit_kind, "t_in": h.t_in, "t_chunk": h.t_chunk, "t_block": h.t_block}


def primary_ray(cam, W, H, px, py, map_dim, epsilon=0.001):
    cam = np.asarray(cam, dtype=CAMERA_DTYPE)
    o = np.zeros(3, np.float32); d = np.zeros(3, np.float32); s = np.zeros(3, np.float32)
    thf = np.tan(np.float32(cam["fov"]) / np.float32(2.0), dtype=np.float32)
    lib().orc_primary_ray(cam.ctypes.data, float(thf), W, H, px, py, map_dim, float(np.float32(epsilon)), o.ctypes.data, d.ctypes.data, s.ctypes.data)
    return o, d, s


def primary(world, cam, W, H, prm=None, want_hits=True):
    prm = prm or params(world.dim)
    cam = np.asarray(cam, dtype=CAMERA_DTYPE)
    albedo = np.empty((H, W), np.uint32)
    normal = np.empty((H, W), np.uint32)
    position = np.empty((H, W, 4), np.float32)
    hits = np.empty((H, W), HIT_DTYPE) if want_hits else None
    cnt = _Counters()
    lib().orc_primary(ctypes.byref(world._c), cam.ctypes.data, ctypes.byref(prm), W, H, albedo.ctypes.data, normal.ctypes.data,
                      position.ctypes.data, hits.ctypes.data if want_hits else None, ctypes.byref(cnt))
    return {"albedo": albedo, "normal": normal, "position": position, "hits": hits, "counters": cnt.as_dict()}


def primary_pixels(world, cam, W, H, xs, ys, prm=None):
    """primary pass of the listed pixels of a W x H frame (sampled parity of frames too large for the CPU in test time);
    every output is indexed by sample.  The shadow pass of the same samples is secondary() on the [1, n] arrays."""
    prm = prm or params(world.dim)
    cam = np.asarray(cam, dtype=CAMERA_DTYPE)
    xs = np.ascontiguousarray(xs, dtype=np.uint32)
    ys = np.ascontiguousarray(ys, dtype=np.uint32)
    n = xs.size
    albedo = np.empty(n, np.uint32)
    normal = np.empty(n, np.uint32)
    position = np.empty((n, 4), np.float32)
    hits = np.empty(n, HIT_DTYPE)
    lib().orc_primary_pixels(ctypes.byref(world._c), cam.ctypes.data, ctypes.byref(prm), W, H, n, xs.ctypes.data, ys.ctypes.data,
                             albedo.ctypes.data, normal.ctypes.data, position.ctypes.data, hits.ctypes.data)
    return {"albedo": albedo, "normal": normal, "position": position, "hits": hits}


def secondary(world, normal, position, prm=None):
    prm = prm or params(world.dim)
    H, W = normal.shape
    illum = np.empty((H, W), np.uint32)
    cnt = _Counters()
    lib().orc_secondary(ctypes.byref(world._c), ctypes.byref(prm), W, H, np.ascontiguousarray(normal).ctypes.data,
                        np.ascontiguousarray(position).ctypes.data, illum.ctypes.data, ctypes.byref(cnt))
    return {"illumination": illum, "counters": cnt.as_dict()}


def blit(albedo, normal, position, illum):
    H, W = albedo.shape
    frame = np.empty((H, W), np.uint32)
    lib().orc_blit(W, H, np.ascontiguousarray(albedo).ctypes.data, np.ascontiguousarray(normal).ctypes.data,
                   np.ascontiguousarray(position).ctypes.data, np.ascontiguousarray(illum).ctypes.data, frame.ctypes.data)
    return frame


def render(world, cam, W, H, prm=None, want_hits=True):
    """The reference frame (src/game.zig:244-255): primary, secondary, blit."""
    p = primary(world, cam, W, H, prm, want_hits)
    s = secondary(world, p["normal"], p["position"], prm)
    f = blit(p["albedo"], p["normal"], p["position"], s["illumination"])
    return {**p, "illumination": s["illumination"], "frame": f, "primary_counters": p["counters"], "secondary_counters": s["counters"]}


def sky_dome2(rd):
    rd = np.asarray(rd, dtype=np.float32)
    out = np.zeros(3, np.float32)
    lib().orc_sky_dome2(rd.ctypes.data, out.ctypes.data)
    return out


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def algorithmic_bytes(counters, pixels, which):
    """bytes the reference's own access pattern moves for one pass."""
    trav = 4 * (counters["t_in"] + counters["t_chunk"] + counters["t_block"])
    if which == "primary":
        return trav + 24 * pixels
    if which == "secondary":
        return trav + 24 * counters["rays"] + 20 * counters["early_out"]
    raise ValueError(which)
