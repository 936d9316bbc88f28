"""Mirror of src/procgen.zig (native implementation: csrc/host/world.cpp, noise.cpp)."""
from . import _native as N
from .gfx import UvtError


def procgen(dim, world, offsetX=0.0, offsetY=0.0, device="auto"):
    """procgen.zig:6: fill `world` (a VoxelBrickmap of the same dim).

    device: "auto" generates on the GPU when the map is attached to a ctx and still empty (the same world byte for byte,
    csrc/procgen.cuh) and falls back to the serial host version otherwise; "host" / "device" force one of them."""
    L = N.load()
    on_ctx = getattr(world, "ctx", None) is not None and not getattr(world.ctx, "is_group", False)
    if device == "device" or (device == "auto" and on_ctx and world.n_bricks == 0):
        rc = L.uvt_procgen_device(world.handle, dim, float(offsetX), float(offsetY))
        if rc == N.UVT_OK:
            return
        if device == "device":
            raise UvtError(rc, "device procgen failed: " + ((L.uvt_last_error(world.ctx.handle) or b"").decode() if on_ctx else "the map is not attached to a ctx"))
    rc = L.uvt_procgen(world.handle, dim, float(offsetX), float(offsetY))
    if rc != N.UVT_OK:
        raise UvtError(rc, "procgen failed")


def height(dim, x, z, offsetX=0.0, offsetY=0.0):
    """terrain height of column (x, z): procgen.zig:23-24."""
    return int(N.load().uvt_procgen_height(dim, x, z, float(offsetX), float(offsetY)))


def noise2(x, y):
    return float(N.load().uvt_noise2_fbm(float(x), float(y)))
