"""Mirror of src/procgen.zig (native implementation: csrc/host/world.cpp, noise.cpp)."""
from . import _native as N
from .gfx import UvtError


def procgen(dim, world, offsetX=0.0, offsetY=0.0):
    """procgen.zig:6: fill `world` (a VoxelBrickmap of the same dim)."""
    rc = N.load().uvt_procgen(world.handle, dim, float(offsetX), float(offsetY))
    if rc != N.UVT_OK:
        raise UvtError(rc, "procgen failed")


def height(dim, x, z, offsetX=0.0, offsetY=0.0):
    """terrain height of column (x, z): procgen.zig:23-24."""
    return int(N.load().uvt_procgen_height(dim, x, z, float(offsetX), float(offsetY)))


def noise2(x, y):
    return float(N.load().uvt_noise2_fbm(float(x), float(y)))
