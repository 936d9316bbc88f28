"""Mirror of src/engine/voxel.zig: Voxel word, VoxelBrickmap(dim, 8), VoxelModelAtlas.

The storage lives in the native library (csrc/host/world.cpp, atlas.cpp); these classes are
the reference-shaped handles over it.  With a ctx the brickmap's storage IS the ctx's pinned
staging and bind() publishes it to the device, like the persistently mapped GL buffers.
"""
import ctypes

import numpy as np

from . import _native as N
from .gfx import UvtError, _check, _check_owner

TY_MASK = 0x0FFFFFFF
SOLID = 0x10000000


def Voxel(ty, is_solid=False):
    """voxel.zig:7-19: packed u32, ty:u28 | is_solid<<28."""
    return (int(ty) & TY_MASK) | (SOLID if is_solid else 0)


Voxel.EMPTY = 0


class VoxelBrickmap:
    def __init__(self, dim, chsize=8, ctx=None):
        if chsize != 8:
            raise UvtError(N.UVT_ERR_INVALID, "the traversal kernels assume CHUNK_DIMENSION 8 (map.glsl:4)")
        self.ctx = ctx
        self.dim = dim
        self._L = N.load()
        h = ctypes.c_void_p()
        if ctx is not None and getattr(ctx, "is_group", False):
            rc = self._L.uvt_brickmap_create_group(ctx.handle, dim, ctypes.byref(h))
        else:
            rc = self._L.uvt_brickmap_create(ctx.handle if ctx else None, dim, ctypes.byref(h))
        _check_owner(ctx, rc)
        self.handle = h

    @classmethod
    def init(cls, dim=512, chsize=8, ctx=None):
        return cls(dim, chsize, ctx)

    @classmethod
    def load(cls, path, ctx=None):
        self = cls.__new__(cls)
        self.ctx, self._L = ctx, N.load()
        h = ctypes.c_void_p()
        rc = self._L.uvt_brickmap_load(ctx.handle if ctx else None, path.encode(), ctypes.byref(h))
        if rc != N.UVT_OK:
            raise UvtError(rc, f"cannot load world dump {path}")
        self.handle = h
        self.dim = self._L.uvt_brickmap_dim(h)
        return self

    def save(self, path):
        rc = self._L.uvt_brickmap_save(self.handle, path.encode())
        if rc != N.UVT_OK:
            raise UvtError(rc, f"cannot write world dump {path}")

    def clear(self, _=0):
        self._L.uvt_brickmap_clear(self.handle)

    def set(self, x, y, z, voxel):
        rc = self._L.uvt_brickmap_set(self.handle, x, y, z, voxel)
        if rc != N.UVT_OK:
            raise UvtError(rc, f"set({x},{y},{z}) failed")

    def get(self, x, y, z):
        return int(self._L.uvt_brickmap_get(self.handle, x, y, z))

    def is_walkable(self, x, y, z):
        return bool(self._L.uvt_brickmap_is_walkable(self.handle, x, y, z))

    @property
    def n_bricks(self):
        return int(self._L.uvt_brickmap_n_bricks(self.handle))

    @property
    def capacity(self):
        return int(self._L.uvt_brickmap_capacity(self.handle))

    def chunks(self):
        """read-only numpy view of the chunk table u32[(dim/8)^3] (no copy; writes go through set())."""
        n = (self.dim // 8) ** 3
        p = self._L.uvt_brickmap_chunks(self.handle)
        a = np.ctypeslib.as_array((ctypes.c_uint32 * n).from_address(p))
        a.flags.writeable = False
        return a

    def bricks(self):
        """read-only numpy view of the committed part of the brick pool u32[n_bricks][512] (no copy)."""
        n = max(self.n_bricks, 1) * 512
        p = self._L.uvt_brickmap_bricks(self.handle)
        a = np.ctypeslib.as_array((ctypes.c_uint32 * n).from_address(p)).reshape(-1, 512)
        a.flags.writeable = False
        return a

    def mark_dirty(self):
        """After writes that bypassed set(): the next bind() republishes the whole map."""
        self._L.uvt_brickmap_mark_dirty(self.handle)

    def bind(self, base_binding=9):
        """voxel.zig:77-80: pool -> binding 9, chunk table -> binding 10.  Publishes what set() wrote since the last
        bind (first bind: everything; clean map: nothing), so calling it every frame like game.zig:236 is cheap."""
        if base_binding != 9:
            raise UvtError(N.UVT_ERR_INVALID, "voxelData/mapData are bindings 9/10 (map.glsl:11-17)")
        if not self.ctx:
            raise UvtError(N.UVT_ERR_INVALID, "brickmap has no ctx to bind to")
        _check_owner(self.ctx, self._L.uvt_brickmap_bind(self.handle))

    def deinit(self):
        if self.handle:
            self._L.uvt_brickmap_destroy(self.handle)
            self.handle = None


class VoxelModelAtlas:
    def __init__(self, ctx=None):
        self.ctx = ctx
        self._L = N.load()
        h = ctypes.c_void_p()
        if ctx is not None and getattr(ctx, "is_group", False):
            rc = self._L.uvt_atlas_create_group(ctx.handle, ctypes.byref(h))
        else:
            rc = self._L.uvt_atlas_create(ctx.handle if ctx else None, ctypes.byref(h))
        if rc != N.UVT_OK:
            raise UvtError(rc, "atlas create failed")
        self.handle = h

    @classmethod
    def init(cls, ctx=None):
        return cls(ctx)

    @property
    def current_index(self):
        return int(self._L.uvt_atlas_current_index(self.handle))

    def load_block_model(self, model):
        """voxel.zig:115-127. `model` is a path, or bytes of a .vox file."""
        if isinstance(model, (bytes, bytearray)):
            buf = (ctypes.c_uint8 * len(model)).from_buffer_copy(model)
            rc = self._L.uvt_atlas_load_block_model_mem(self.handle, buf, len(model))
        else:
            rc = self._L.uvt_atlas_load_block_model(self.handle, str(model).encode())
        if rc != N.UVT_OK:
            detail = self._L.uvt_vox_error().decode()
            if self.ctx and rc in (N.UVT_ERR_CUDA, N.UVT_ERR_INVALID):
                detail = detail or self.ctx.L.uvt_last_error(self.ctx.handle).decode()
            raise UvtError(rc, f"load_block_model({model if not isinstance(model, (bytes, bytearray)) else '<bytes>'}): {detail}")

    def append_model(self, texels):
        t = np.ascontiguousarray(texels, dtype=np.uint32).reshape(512)
        rc = self._L.uvt_atlas_append_model(self.handle, t.ctypes.data)
        if rc != N.UVT_OK:
            raise UvtError(rc, "append_model failed")

    def models(self):
        """[n,512] texels of every loaded model (x + 8y + 64z)."""
        out = np.zeros((self.current_index, 512), dtype=np.uint32)
        for i in range(self.current_index):
            self._L.uvt_atlas_get_model(self.handle, i, out[i].ctypes.data)
        return out

    def bind(self, idx=6):
        if idx != 6:
            raise UvtError(N.UVT_ERR_INVALID, "the model atlas is image unit 6 (map.glsl:19)")

    def deinit(self):
        if self.handle:
            self._L.uvt_atlas_destroy(self.handle)
            self.handle = None


def parse_vox(data):
    """zvox.VoxFile.from_reader stand-in: returns (models=[(size, voxels[n,4])], palette[256])."""
    L = N.load()
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data)
    h = ctypes.c_void_p()
    rc = L.uvt_vox_parse(buf, len(data), ctypes.byref(h))
    if rc != N.UVT_OK:
        raise UvtError(rc, L.uvt_vox_error().decode())
    try:
        models = []
        for m in range(L.uvt_vox_n_models(h)):
            size = (ctypes.c_uint32 * 3)()
            L.uvt_vox_model_size(h, m, ctypes.byref(size))
            n = L.uvt_vox_model_n_voxels(h, m)
            vox = np.zeros((n, 4), dtype=np.uint8)
            if n:
                ctypes.memmove(vox.ctypes.data, L.uvt_vox_model_voxels(h, m), n * 4)
            models.append((tuple(size), vox))
        pal = np.zeros(256, dtype=np.uint32)
        ctypes.memmove(pal.ctypes.data, L.uvt_vox_palette(h), 1024)
        return models, pal
    finally:
        L.uvt_vox_free(h)


def load_model(data, size):
    """The call src/game.zig:114 keeps commented out — `models.load_model("assets/chicken.vox", allocator, 32)` — for
    the entity model of uvt_entity_model_upload: the first model of a .vox file as [size^3] RGBA8 texels indexed
    x + size * (y + size * z), converted like load_single_block_model (voxel.zig:104-108: .vox z is up, so y and z
    swap; colour = palette[index - 1]).  `data` is a path or the bytes of the file."""
    if not isinstance(data, (bytes, bytearray)):
        with open(data, "rb") as f:
            data = f.read()
    models, pal = parse_vox(data)
    if not models:
        raise UvtError(N.UVT_ERR_FORMAT, "no model in the .vox file")
    (sx, sy, sz), vox = models[0]
    if max(sx, sy, sz) > size:
        raise UvtError(N.UVT_ERR_FORMAT, f"model is {sx}x{sy}x{sz}, larger than {size}^3")
    out = np.zeros(size ** 3, dtype=np.uint32)
    v = vox.astype(np.int64)
    out[v[:, 0] + size * (v[:, 2] + size * v[:, 1])] = pal[(v[:, 3] - 1) & 255]
    return out
