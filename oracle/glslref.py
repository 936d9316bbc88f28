"""ctypes front end of oracle/_ref/libglslref.so — the reference's OWN shader text (assets/shaders/*.glsl) compiled for the
CPU against a GLSL-in-C++ shim (oracle/glsl_ref/).  TEST INFRASTRUCTURE ONLY: it exists to pin oracle/oracle.c, and
through it the CUDA path, against the reference's source instead of against a restatement by the same author.

`variant="entities"` loads the library built from the text with map.glsl:199 deleted and primary.comp.glsl:47-54
uncommented; the default is the unmodified text.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE = os.environ.get("UVT_REFERENCE", "/root/reference")
_libs = {}


def _path(variant):
    return os.path.join(_REF_DIR, "libglslref_entities.so" if variant == "entities" else "libglslref.so")


def build(force=False):
    """(Re)build oracle/_ref from the reference checkout when it is present; a no-op otherwise (the GPU box only has the
    prebuilt libraries)."""
    if not os.path.isdir(os.path.join(REFERENCE, "assets", "shaders")):
        return os.path.exists(_path(""))
    srcs = [os.path.join(_HERE, "glsl_ref", f) for f in ("glsl_shim.h", "harness.cpp", "translate.py")]
    srcs += [os.path.join(REFERENCE, "assets", "shaders", f) for f in os.listdir(os.path.join(REFERENCE, "assets", "shaders"))]
    libs = [_path(""), _path("entities")]
    if not force and all(os.path.exists(p) for p in libs) and min(os.path.getmtime(p) for p in libs) >= max(os.path.getmtime(s) for s in srcs):
        return True
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", _HERE, "ref", f"REFERENCE={REFERENCE}"], check=True, env=env, stdout=subprocess.DEVNULL)
    return True


def available(variant=""):
    try:
        build()
    except Exception:
        pass
    return os.path.exists(_path(variant))


def lib(variant=""):
    if variant not in _libs:
        if not available(variant):
            raise RuntimeError("oracle/_ref is not built and the reference checkout is absent")
        L = ctypes.CDLL(_path(variant))
        v, u32, f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_float
        L.ref_info.restype = ctypes.c_char_p
        L.ref_primary.argtypes = [u32, v, v, v, v, u32, u32, v, v, v]
        L.ref_secondary.argtypes = [u32, v, v, v, u32, u32, v, v, v, v]
        L.ref_blit.argtypes = [u32, u32, v, v, v, v, v]
        L.ref_trace_map.argtypes = [u32, v, v, v, v, v, ctypes.c_int, v, v, v]
        L.ref_trace_entities.argtypes = [v, v, v, f32, v, v, v]
        L.ref_sky_dome2.argtypes = [v, v]
        L.ref_num_threads.restype = ctypes.c_int
        L.ref_set_num_threads.argtypes = [ctypes.c_int]
        _libs[variant] = L
    return _libs[variant]


def render(world, cam, W, H, variant=""):
    """The reference frame (src/game.zig:244-255) by the reference's shader text: G-buffer images + final frame.
    `world` is an oracle.World (chunk table, brick pool, 256^3 atlas); arrays are [H][W], row 0 = bottom row."""
    L = lib(variant)
    cam = np.ascontiguousarray(cam)
    assert cam.nbytes == 96
    albedo = np.zeros((H, W), np.uint32)
    normal = np.zeros((H, W), np.uint32)
    position = np.zeros((H, W, 4), np.float32)
    illum = np.zeros((H, W), np.uint32)
    frame = np.zeros((H, W), np.uint32)
    args = (world.dim, world.chunks.ctypes.data, world.bricks.ctypes.data, world.atlas.ctypes.data)
    L.ref_primary(*args, cam.ctypes.data, W, H, albedo.ctypes.data, normal.ctypes.data, position.ctypes.data)
    L.ref_secondary(*args, W, H, albedo.ctypes.data, normal.ctypes.data, position.ctypes.data, illum.ctypes.data)
    L.ref_blit(W, H, albedo.ctypes.data, normal.ctypes.data, position.ctypes.data, illum.ctypes.data, frame.ctypes.data)
    return {"albedo": albedo, "normal": normal, "position": position, "illumination": illum, "frame": frame}


def primary(world, cam, W, H, variant=""):
    """primary.comp.glsl alone (BASELINE config 2 is primary rays only)."""
    L = lib(variant)
    cam = np.ascontiguousarray(cam)
    albedo = np.zeros((H, W), np.uint32)
    normal = np.zeros((H, W), np.uint32)
    position = np.zeros((H, W, 4), np.float32)
    L.ref_primary(world.dim, world.chunks.ctypes.data, world.bricks.ctypes.data, world.atlas.ctypes.data, cam.ctypes.data, W, H,
                  albedo.ctypes.data, normal.ctypes.data, position.ctypes.data)
    return {"albedo": albedo, "normal": normal, "position": position}


def shadow_rays(position):
    """pixels whose shadow ray is traced: secondary.comp.glsl:26-29 leaves where a position component is negative"""
    return int((position[..., :3] >= 0).all(-1).sum())


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    for v in list(_libs) or [""]:
        lib(v).ref_set_num_threads(int(n))


def trace_map(world, origin, direction, max_steps, variant=""):
    o = np.asarray(origin, np.float32)
    d = np.asarray(direction, np.float32)
    data = ctypes.c_uint32()
    hp = np.zeros(3, np.float32)
    n = np.zeros(3, np.float32)
    lib(variant).ref_trace_map(world.dim, world.chunks.ctypes.data, world.bricks.ctypes.data, world.atlas.ctypes.data, o.ctypes.data, d.ctypes.data,
                               int(max_steps), ctypes.byref(data), hp.ctypes.data, n.ctypes.data)
    return {"data": data.value, "hit_pos": tuple(float(x) for x in hp), "normal": tuple(float(x) for x in n)}


def trace_entities(world, origin, direction, max_distance, variant=""):
    o = np.asarray(origin, np.float32)
    d = np.asarray(direction, np.float32)
    data = ctypes.c_uint32()
    hp = np.zeros(3, np.float32)
    n = np.zeros(3, np.float32)
    lib(variant).ref_trace_entities(world.atlas.ctypes.data, o.ctypes.data, d.ctypes.data, float(np.float32(max_distance)), ctypes.byref(data), hp.ctypes.data, n.ctypes.data)
    return {"data": data.value, "hit_pos": tuple(float(x) for x in hp), "normal": tuple(float(x) for x in n)}


def sky_dome2(rd):
    rd = np.asarray(rd, np.float32)
    out = np.zeros(4, np.float32)
    lib().ref_sky_dome2(rd.ctypes.data, out.ctypes.data)
    return out
