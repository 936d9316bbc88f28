"""The benchmark inputs of BASELINE.json / SURVEY §8d: worlds W1 / W4, cameras K0 / K1 and the C5 pose sweep."""
import numpy as np

from . import _native as N
from . import procgen, voxel

LCG_SEED = 0x46AE4F


def camera(pos, mat=None, fov=np.pi / 2):
    cam = np.zeros((), dtype=N.CAMERA_DTYPE)
    cam["cam_pos"][:3] = pos
    cam["cam_mat"] = np.eye(4, dtype=np.float32).reshape(16) if mat is None else np.asarray(mat, np.float32).reshape(16)
    cam["fov"] = np.float32(fov)
    return cam


def pitch_yaw(pitch, yaw):
    import ctypes
    m = (ctypes.c_float * 16)()
    N.load().uvt_mat_from_pitch_yaw(float(pitch), float(yaw), ctypes.byref(m))
    return np.array(m, dtype=np.float32)


def camera_k0(dim=512):
    """spawn (256,22,256) + (0,3,0), identity, fov pi/2 (game.zig:40,212); for other dims: map centre at terrain height + 3."""
    if dim == 512:
        return camera((256.0, 25.0, 256.0))
    c = dim // 2
    return camera((float(c), float(max(procgen.height(dim, c, c), 16) + 3), float(c)))


def camera_k1(dim=512):
    """pitched/yawed view over the hills (trees + shadows in view).  The same block coordinates for every world
    size: procgen only plants trees for x, z < 500 (procgen.zig:47), so the W4 camera stays in that corner."""
    base = max(procgen.height(dim, 200, 140), 16)
    return camera((200.0, float(base + 28), 140.0), pitch_yaw(0.35, 0.6))


def sweep_poses(dim, n=256, seed=LCG_SEED):
    """C5: position uniform in the interior at terrain height + U[3,40], yaw U[0,2pi), pitch U[-0.6,0.6];
    PRNG = the reference LCG (util.zig:33-45) so the list is reproducible in any language."""
    state = [seed]

    def u():
        state[0] = (state[0] * 1103515245 + 12345) & 0xFFFFFFFF
        return state[0] / 4294967296.0

    cams = np.zeros(n, dtype=N.CAMERA_DTYPE)
    for i in range(n):
        x = 0.05 * dim + u() * 0.9 * dim
        z = 0.05 * dim + u() * 0.9 * dim
        y = max(procgen.height(dim, int(x), int(z)), 16) + 3.0 + u() * 37.0
        yaw = u() * 2.0 * np.pi
        pitch = -0.6 + u() * 1.2
        cams[i] = camera((x, y, z), pitch_yaw(pitch, yaw))
    return cams


def build_world(ctx, dim, models):
    """procgen(dim) into the ctx's pinned staging + the block models, committed through the C ABI."""
    bm = voxel.VoxelBrickmap.init(dim, 8, ctx)
    procgen.procgen(dim, bm)
    atlas = voxel.VoxelModelAtlas.init(ctx)
    for m in np.asarray(models, dtype=np.uint32).reshape(-1, 512):
        atlas.append_model(m)
    bm.bind(9)
    return bm, atlas
