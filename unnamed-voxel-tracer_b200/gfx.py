"""Host-side mirror of the reference's `gfx` wrapper API over the C ABI.

Same names, argument meaning and error behaviour as src/engine/graphics/*.zig so code
written against the reference reads the same here:

    ComputePipeline.init / bind / dispatch / deinit        shader.zig:97-122
    RasterPipeline.init / bind / draw / deinit              shader.zig:125-153
    PersistentMappedBuffer(UniformData).deref / bind        buffer.zig:80-134
    GBuffer.init / resize / bind_images / bind_textures     gbuffer.zig:3-53
    Camera / Camera.UniformData                             camera.zig:5-43

Errors: the reference returns Zig error unions at init and @panic's elsewhere; every call
here raises UvtError (carrying the C status and uvt_last_error text) instead.
"""
import ctypes

import numpy as np

from . import _native as N


class UvtError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"uvt status {status}: {message}")
        self.status = status
        self.message = message


class ShaderCompilationError(UvtError):
    """compileShader failure (shader.zig:64-67): the kernel image could not be loaded."""


def _check(ctx_handle, rc):
    if rc != N.UVT_OK:
        msg = N.load().uvt_last_error(ctx_handle)
        raise UvtError(rc, msg.decode() if msg else "")
    return rc


def _check_owner(owner, rc):
    """owner: a Context or a Group (or None)."""
    if rc != N.UVT_OK:
        if owner is not None and getattr(owner, "is_group", False):
            msg = N.load().uvt_group_last_error(owner.handle)
        else:
            msg = N.load().uvt_last_error(owner.handle if owner is not None else None)
        raise UvtError(rc, msg.decode() if msg else "")
    return rc


class Context:
    """gfx.init (graphics.zig:60-75): one CUDA device + one in-order stream."""

    def __init__(self, device=0, *, map_dim=512, primary_max_steps=192, shadow_max_steps=48, hit_buffer=False,
                 entities=True, layout="compact", dense=True, fused_frame=False):
        L = N.load()
        p = N.Params()
        L.uvt_default_params(ctypes.byref(p))
        p.map_dim = map_dim
        p.primary_max_steps = primary_max_steps
        p.shadow_max_steps = shadow_max_steps
        p.flags = ((N.UVT_FLAG_HIT_BUFFER if hit_buffer else 0) | (N.UVT_FLAG_ENTITIES if entities else 0) |
                   (0 if dense else N.UVT_FLAG_NO_DENSE) | (N.UVT_FLAG_FUSED_FRAME if fused_frame else 0))
        p.layout = N.UVT_LAYOUT_COMPACT if layout == "compact" else N.UVT_LAYOUT_REFERENCE
        h = ctypes.c_void_p()
        rc = L.uvt_create(ctypes.byref(p), int(device), ctypes.byref(h))
        if rc != N.UVT_OK:
            msg = L.uvt_last_error(None)
            raise UvtError(rc, msg.decode() if msg else "")
        self.handle = h
        self.L = L
        self.device = int(device)
        self._pinned = []

    # -- lifetime
    def close(self):
        if getattr(self, "handle", None):
            self.L.uvt_destroy(self.handle)  # also frees the pinned buffers handed out by pinned_empty()
            self._pinned = []
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        return _check(self.handle, rc)

    # -- thin wrappers over the per-frame C calls
    def set_stream(self, cuda_stream):
        self.check(self.L.uvt_set_stream(self.handle, ctypes.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def set_layout(self, layout):
        self.check(self.L.uvt_set_layout(self.handle, N.UVT_LAYOUT_COMPACT if layout == "compact" else N.UVT_LAYOUT_REFERENCE))

    def set_scheduler(self, scheduler):
        """'tile' (one pixel per thread, the default) or 'pool' (per-CTA ray pool with phase-wise compaction, opt-in)."""
        self.check(self.L.uvt_set_scheduler(self.handle, N.UVT_SCHED_POOL if scheduler == "pool" else N.UVT_SCHED_TILE))

    def effective_layout(self):
        return "compact" if self.L.uvt_effective_layout(self.handle) == N.UVT_LAYOUT_COMPACT else "reference"

    def set_max_steps(self, primary, shadow):
        self.check(self.L.uvt_set_max_steps(self.handle, primary, shadow))

    def set_camera(self, cam):
        cam = np.ascontiguousarray(cam, dtype=N.CAMERA_DTYPE)
        if cam.size == 1:
            self.check(self.L.uvt_set_camera(self.handle, cam.ctypes.data))
        else:
            self.check(self.L.uvt_set_cameras(self.handle, cam.ctypes.data, int(cam.size)))

    def resize(self, w, h):
        self.check(self.L.uvt_resize(self.handle, int(w), int(h)))
        self.width, self.height = int(w), int(h)

    def set_partition(self, band_rows, n_parts, part):
        self.check(self.L.uvt_set_partition(self.handle, band_rows, n_parts, part))

    def local_rows(self):
        r = ctypes.c_uint32()
        self.check(self.L.uvt_local_rows(self.handle, ctypes.byref(r)))
        return r.value

    def dispatch_primary(self):
        self.check(self.L.uvt_dispatch_primary(self.handle))

    def dispatch_secondary(self):
        self.check(self.L.uvt_dispatch_secondary(self.handle))

    def shade(self):
        self.check(self.L.uvt_shade(self.handle))

    def dispatch_secondary_shade(self):
        """dispatch_secondary() + shade() in one launch."""
        self.check(self.L.uvt_dispatch_secondary_shade(self.handle))

    def dispatch_frame(self):
        self.check(self.L.uvt_dispatch_frame(self.handle))

    def world_set_voxel(self, x, y, z, voxel):
        """map_setVoxel (map.glsl:49-55): write + publish one block; False where the chunk holds no brick."""
        w = ctypes.c_int(0)
        self.check(self.L.uvt_world_set_voxel(self.handle, x, y, z, voxel, ctypes.byref(w)))
        return bool(w.value)

    def world_layout_checksum(self):
        out = (ctypes.c_uint64 * 4)()
        self.check(self.L.uvt_world_layout_checksum(self.handle, ctypes.byref(out)))
        d, b, q, m = (int(v) for v in out)
        return d, b, q, m & 0xFFFFFFFF, m >> 32   # dense, bricks, clear4, y_clear, n_materials

    def sync(self):
        self.check(self.L.uvt_sync(self.handle))

    def set_frame_chunks(self, n):
        """dispatch_frame in n row chunks alternating between two streams (hides pass tails when the frame share is small)."""
        self.check(self.L.uvt_set_frame_chunks(self.handle, int(n)))

    def set_entity_mode(self, mode):
        """traceEntities (map.glsl:172-248): "boxes" = the reference as it runs (returns at :199), "models" = the sub-model
        DDA behind that return + the primary-pass composite of primary.comp.glsl:45-54 (turns the hit buffer on)."""
        self.check(self.L.uvt_set_entity_mode(self.handle, {"boxes": N.UVT_ENTITY_BOXES, "models": N.UVT_ENTITY_MODELS}[mode]))

    def set_entities(self, positions=None):
        """`positions[]` of map.glsl:173-179 (low box corners, blocks); None restores the five literals."""
        if positions is None:
            self.check(self.L.uvt_set_entities(self.handle, None, 0))
            return
        p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        self.check(self.L.uvt_set_entities(self.handle, p.ctypes.data, len(p)))

    def entity_model_upload(self, model=None, size=8, max_steps=0):
        """size^3 RGBA8 texels [z][y][x] (8, 16 or 32: chicken.vox); None = texels [0,8)^3 of the atlas (map.glsl:218)."""
        if model is None:
            self.check(self.L.uvt_entity_model_upload(self.handle, 8, None, int(max_steps)))
            return
        m = np.ascontiguousarray(model, dtype=np.uint32).reshape(-1)
        if m.size != size ** 3:
            raise ValueError("model must hold size^3 texels")
        self.check(self.L.uvt_entity_model_upload(self.handle, int(size), m.ctypes.data, int(max_steps)))

    def pick(self):
        out = np.zeros((), dtype=N.HIT_DTYPE)
        self.check(self.L.uvt_pick(self.handle, out.ctypes.data))
        return out

    def buffer_bytes(self, kind):
        return int(self.L.uvt_buffer_bytes(self.handle, kind))

    def device_ptr(self, kind):
        p = ctypes.c_void_p()
        self.check(self.L.uvt_device_ptr(self.handle, kind, ctypes.byref(p)))
        return p.value

    def bind_frame_target(self, dptr, global_rows=True):
        self.check(self.L.uvt_bind_frame_target(self.handle, ctypes.c_void_p(dptr or 0), 0, 1 if global_rows else 0))

    def deinterleave(self, gathered_ptr, frame_ptr, rows_per_part):
        self.check(self.L.uvt_deinterleave(self.handle, ctypes.c_void_p(gathered_ptr), ctypes.c_void_p(frame_ptr), rows_per_part))

    def shared_frame_create(self):
        """Presenting rank: allocate the full WxH frame and return (device pointer, 64-byte CUDA IPC handle)."""
        p = ctypes.c_void_p()
        h = (ctypes.c_ubyte * 64)()
        self.check(self.L.uvt_shared_frame_create(self.handle, ctypes.byref(p), h))
        return p.value, bytes(h)

    def shared_frame_open(self, handle):
        """Other ranks: map the presenting rank's frame into this process (peer access over NVLink)."""
        p = ctypes.c_void_p()
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        self.check(self.L.uvt_shared_frame_open(self.handle, buf, ctypes.byref(p)))
        return p.value

    def shared_frame_close(self, dptr):
        self.check(self.L.uvt_shared_frame_close(self.handle, ctypes.c_void_p(dptr)))

    def read_device(self, dptr, out):
        self.check(self.L.uvt_read_device(self.handle, ctypes.c_void_p(dptr), out.ctypes.data, out.nbytes))
        return out

    _KINDS = {"albedo": (N.UVT_BUF_ALBEDO, np.uint32, ()), "normal": (N.UVT_BUF_NORMAL, np.uint32, ()),
              "position": (N.UVT_BUF_POSITION, np.float32, (4,)), "illumination": (N.UVT_BUF_ILLUMINATION, np.uint32, ()),
              "frame": (N.UVT_BUF_FRAME, np.uint32, ()), "hit": (N.UVT_BUF_HIT, N.HIT_DTYPE, ())}

    def readback(self, name, out=None):
        """Copy a G-buffer image to the host: array [layers?, rows, W(, 4)], row 0 = bottom image row."""
        kind, dtype, tail = self._KINDS[name]
        nbytes = self.buffer_bytes(kind)
        if nbytes == 0:
            raise UvtError(N.UVT_ERR_INVALID, f"buffer '{name}' is not allocated")
        if out is None:
            out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        self.check(self.L.uvt_readback(self.handle, kind, out.ctypes.data, nbytes))
        rows = self.local_rows()
        shape = (rows, self.width) + tail
        per_layer = int(np.prod(shape))
        layers = out.size // per_layer
        return out.reshape((layers,) + shape) if layers > 1 else out.reshape(shape)

    def pinned_empty(self, nbytes, dtype=np.uint8):
        """numpy view over pinned host memory owned by the ctx: uvt_destroy frees whatever free_pinned() has not.
        The view must not be used after close()."""
        p = ctypes.c_void_p()
        self.check(self.L.uvt_alloc_pinned(self.handle, nbytes, ctypes.byref(p)))
        self._pinned.append(p)
        buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype)

    def free_pinned(self, arr):
        """Release a pinned_empty() buffer early (the view is dead afterwards)."""
        addr = arr.ctypes.data
        for i, p in enumerate(self._pinned):
            if p.value == addr:
                self.check(self.L.uvt_free_pinned(self.handle, p))
                del self._pinned[i]
                return
        raise UvtError(N.UVT_ERR_INVALID, "not a pinned_empty() buffer of this ctx")

    def readback_into(self, name, pinned):
        kind, _, _ = self._KINDS[name]
        self.check(self.L.uvt_readback(self.handle, kind, pinned.ctypes.data, pinned.nbytes))

    def readback_async(self, name, pinned):
        """Pipelined readback into pinned host memory (see uvt_readback_async); pair with readback_wait()."""
        kind, _, _ = self._KINDS[name]
        self.check(self.L.uvt_readback_async(self.handle, kind, pinned.ctypes.data, pinned.nbytes))

    @staticmethod
    def nccl_unique_id():
        """128-byte NCCL unique id (rank 0 creates it, the host hands it to every rank)."""
        buf = (ctypes.c_ubyte * 128)()
        rc = N.load().uvt_nccl_unique_id(buf)
        if rc != N.UVT_OK:
            raise UvtError(rc, N.load().uvt_last_error(None).decode())
        return bytes(buf)

    def nccl_init(self, unique_id, n_ranks, rank):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id)
        self.check(self.L.uvt_nccl_init(self.handle, buf, n_ranks, rank))

    def nccl_shutdown(self):
        self.check(self.L.uvt_nccl_shutdown(self.handle))

    def dispatch_frame_nccl(self, full_frame_ptr=None, n_groups=4):
        """The tiled frame with the NCCL band exchange overlapped with traversal (uvt_dispatch_frame_nccl)."""
        self.check(self.L.uvt_dispatch_frame_nccl(self.handle, ctypes.c_void_p(full_frame_ptr or 0), n_groups))

    def readback_bands_async(self, host_frame):
        """This ctx's frame bands into their rows of a full W x H uint32 host frame (shared by all ranks of a tiled frame)."""
        self.check(self.L.uvt_readback_bands_async(self.handle, host_frame.ctypes.data, host_frame.nbytes))

    def host_register(self, arr):
        """Page-lock caller memory (numpy array over e.g. a shared-memory mapping) for asynchronous device-to-host copies."""
        self.check(self.L.uvt_host_register(self.handle, arr.ctypes.data, arr.nbytes))

    def host_unregister(self, arr):
        self.check(self.L.uvt_host_unregister(self.handle, arr.ctypes.data))

    def readback_wait(self):
        self.check(self.L.uvt_readback_wait(self.handle))

    def count_pass(self, which):
        c = N.Counters()
        self.check(self.L.uvt_count_pass(self.handle, 0 if which == "primary" else 1, ctypes.byref(c)))
        return c.as_dict()

    def fetch_stats(self, which):
        """Fast-path statistics of one pass: {'rays', 'lookups'} — trips that actually fetched world data."""
        c = N.Counters()
        self.check(self.L.uvt_count_pass(self.handle, 2 if which == "primary" else 3, ctypes.byref(c)))
        return {"rays": int(c.rays), "lookups": int(c.t_in), "hits": int(c.hits)}

    def enable_timing(self, on=True):
        self.check(self.L.uvt_enable_timing(self.handle, 1 if on else 0))

    def last_pass_ms(self, which):
        idx = {"primary": 0, "secondary": 1, "shade": 2, "frame": 3}[which]
        ms = ctypes.c_float()
        self.check(self.L.uvt_last_pass_ms(self.handle, idx, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.L.uvt_launch_count(self.handle))

    def measure_l2_read_gbps(self, nbytes=32 << 20, repeats=50):
        g = ctypes.c_float()
        self.check(self.L.uvt_measure_l2_read_gbps(self.handle, nbytes, repeats, ctypes.byref(g)))
        return g.value

    def measure_hbm_copy_gbps(self, nbytes=1 << 30, repeats=10):
        g = ctypes.c_float()
        self.check(self.L.uvt_measure_hbm_copy_gbps(self.handle, nbytes, repeats, ctypes.byref(g)))
        return g.value


def init(device=0, **kw):
    """gfx.init(window) (graphics.zig:60): returns the context every wrapper below hangs off."""
    return Context(device, **kw)



class Group:
    """Several GPUs behind one handle in ONE process (uvt_group, include/uvt.h): world and atlas replicated, the frame
    cut into interleaved 16-row bands, every member storing its bands into member 0's frame over NVLink."""

    is_group = True

    def __init__(self, devices, *, map_dim=512, primary_max_steps=192, shadow_max_steps=48, hit_buffer=False, entities=True):
        L = N.load()
        p = N.Params()
        L.uvt_default_params(ctypes.byref(p))
        p.map_dim = map_dim
        p.primary_max_steps = primary_max_steps
        p.shadow_max_steps = shadow_max_steps
        p.flags = (N.UVT_FLAG_HIT_BUFFER if hit_buffer else 0) | (N.UVT_FLAG_ENTITIES if entities else 0)
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p()
        rc = L.uvt_group_create(ctypes.byref(p), devs, len(devices), ctypes.byref(h))
        if rc != N.UVT_OK:
            msg = L.uvt_group_last_error(None)
            raise UvtError(rc, msg.decode() if msg else "")
        self.handle, self.L, self.devices = h, L, list(devices)
        self.W = self.H = 0

    def close(self):
        if getattr(self, "handle", None):
            self.L.uvt_group_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def check(self, rc):
        return _check_owner(self, rc)

    @property
    def size(self):
        return int(self.L.uvt_group_size(self.handle))

    def member_handle(self, i):
        return ctypes.c_void_p(self.L.uvt_group_member(self.handle, i))

    def set_camera(self, cam):
        buf = np.ascontiguousarray(cam)
        assert buf.nbytes == 96
        self.check(self.L.uvt_group_set_camera(self.handle, buf.ctypes.data))

    def set_entity_mode(self, mode):
        self.check(self.L.uvt_group_set_entity_mode(self.handle, {"boxes": N.UVT_ENTITY_BOXES, "models": N.UVT_ENTITY_MODELS}[mode]))

    def set_entities(self, positions=None):
        if positions is None:
            self.check(self.L.uvt_group_set_entities(self.handle, None, 0))
            return
        p = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        self.check(self.L.uvt_group_set_entities(self.handle, p.ctypes.data, len(p)))

    def entity_model_upload(self, model=None, size=8, max_steps=0):
        if model is None:
            self.check(self.L.uvt_group_entity_model_upload(self.handle, 8, None, int(max_steps)))
            return
        m = np.ascontiguousarray(model, dtype=np.uint32).reshape(-1)
        self.check(self.L.uvt_group_entity_model_upload(self.handle, int(size), m.ctypes.data, int(max_steps)))

    def resize(self, w, h):
        self.check(self.L.uvt_group_resize(self.handle, w, h))
        self.W, self.H = w, h

    def dispatch_frame(self):
        self.check(self.L.uvt_group_dispatch_frame(self.handle))

    def sync(self):
        self.check(self.L.uvt_group_sync(self.handle))

    def readback_frame(self, out=None):
        if out is None:
            out = np.empty((self.H, self.W), dtype=np.uint32)
        self.check(self.L.uvt_group_readback_frame(self.handle, out.ctypes.data, out.nbytes))
        return out

    def count_pass(self, which):
        c = N.Counters()
        self.check(self.L.uvt_group_count_pass(self.handle, {"primary": 0, "secondary": 1}[which], ctypes.byref(c)))
        return c.as_dict()


class ComputePipeline:
    """shader.zig:97-122.  `file` selects the pass by the reference shader's file name."""

    _BY_FILE = {"primary.comp.glsl": N.UVT_PIPELINE_PRIMARY, "secondary.comp.glsl": N.UVT_PIPELINE_SECONDARY,
                "terrain_edit.comp.glsl": N.UVT_PIPELINE_EDIT}

    def __init__(self, ctx, handle, kind):
        self.ctx, self.pipeline, self.kind = ctx, handle, kind

    @classmethod
    def init(cls, ctx, file):
        base = file.replace("\\", "/").rsplit("/", 1)[-1]
        if base not in cls._BY_FILE:
            raise ShaderCompilationError(N.UVT_ERR_INVALID, f"no precompiled kernel stands in for '{file}'")
        h = ctypes.c_void_p()
        rc = ctx.L.uvt_pipeline_create(ctx.handle, cls._BY_FILE[base], ctypes.byref(h))
        if rc != N.UVT_OK:
            raise ShaderCompilationError(rc, ctx.L.uvt_last_error(ctx.handle).decode())
        return cls(ctx, h, cls._BY_FILE[base])

    def bind(self):  # gl.useProgram: nothing to select, dispatch names the pass
        pass

    def dispatch(self, x, y, z):
        self.ctx.check(self.ctx.L.uvt_pipeline_dispatch(self.pipeline, x, y, z))

    def deinit(self):
        if self.pipeline:
            self.ctx.L.uvt_pipeline_destroy(self.pipeline)
            self.pipeline = None


class RasterPipeline:
    """shader.zig:125-153: the full-screen blit."""

    def __init__(self, ctx, handle):
        self.ctx, self.pipeline = ctx, handle

    @classmethod
    def init(cls, ctx, vertex, frag):
        for f, want in ((vertex, "blit.vertex.glsl"), (frag, "blit.fragment.glsl")):
            if f.replace("\\", "/").rsplit("/", 1)[-1] != want:
                raise ShaderCompilationError(N.UVT_ERR_INVALID, f"no precompiled kernel stands in for '{f}'")
        h = ctypes.c_void_p()
        rc = ctx.L.uvt_pipeline_create(ctx.handle, N.UVT_PIPELINE_BLIT, ctypes.byref(h))
        if rc != N.UVT_OK:
            raise ShaderCompilationError(rc, ctx.L.uvt_last_error(ctx.handle).decode())
        return cls(ctx, h)

    def bind(self):
        pass

    def draw(self, max_idx):
        if max_idx != 4:
            raise UvtError(N.UVT_ERR_INVALID, "the blit draws a 4-vertex triangle strip")
        self.ctx.check(self.ctx.L.uvt_pipeline_dispatch(self.pipeline, 1, 1, 1))

    def deinit(self):
        if self.pipeline:
            self.ctx.L.uvt_pipeline_destroy(self.pipeline)
            self.pipeline = None


class Camera:
    """camera.zig:5-43 over the C helpers of uvt_host.h."""

    UniformData = N.CAMERA_DTYPE

    def __init__(self):
        self._s = N.CameraState()
        N.load().uvt_camera_init(ctypes.byref(self._s))

    fov = property(lambda s: s._s.fov)
    pitch = property(lambda s: s._s.pitch)
    yaw = property(lambda s: s._s.yaw)

    def rotate(self, pitch, yaw):
        N.load().uvt_camera_rotate(ctypes.byref(self._s), float(pitch), float(yaw))

    def set_pos(self, pos):
        v = (ctypes.c_float * 4)(*[float(x) for x in pos])
        N.load().uvt_camera_set_pos(ctypes.byref(self._s), ctypes.byref(v))

    def incrementFov(self, increment):
        N.load().uvt_camera_increment_fov(ctypes.byref(self._s), float(increment))

    def camera_mat(self):
        return np.array(self._s.cam_mat, dtype=np.float32).reshape(4, 4)

    def as_uniform_data(self):
        out = np.zeros((), dtype=N.CAMERA_DTYPE)
        N.load().uvt_camera_as_uniform_data(ctypes.byref(self._s), out.ctypes.data)
        return out


class PersistentMappedBuffer:
    """buffer.zig:80-134 for the camera UBO: deref() is host memory the game writes each
    frame (game.zig:224-229); bind(8) publishes it to the device (game.zig:235)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._data = np.zeros((), dtype=N.CAMERA_DTYPE)

    @classmethod
    def init(cls, ctx):
        return cls(ctx)

    def deref(self):
        return self._data

    def bind(self, index):
        if index != 8:
            raise UvtError(N.UVT_ERR_INVALID, "the camera uniform block is binding 8 (camera.glsl:2)")
        self.ctx.set_camera(self._data)


class GBuffer:
    """gbuffer.zig:3-53: albedo RGBA8, normal RGBA8, position RGBA32F, illumination RGBA8."""

    def __init__(self, ctx, width, height):
        self.ctx = ctx
        self.width, self.height = width, height
        ctx.resize(width, height)

    @classmethod
    def init(cls, ctx, width, height):
        return cls(ctx, width, height)

    def resize(self, width, height):
        self.width, self.height = width, height
        self.ctx.resize(width, height)

    def bind_images(self, base):
        if base != 0:
            raise UvtError(N.UVT_ERR_INVALID, "G-buffer images are units 0-3 (primary.comp.glsl:7-9)")

    def bind_textures(self, base):
        if base != 0:
            raise UvtError(N.UVT_ERR_INVALID, "G-buffer samplers are units 0-3 (blit.fragment.glsl:5-8)")

    def deinit(self):
        pass
