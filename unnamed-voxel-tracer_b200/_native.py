"""ctypes declarations for libuvt.so (include/uvt.h + include/uvt_host.h).

The library is built in-tree by build.py.  Importing this module never falls back to
anything: if libuvt.so is missing and cannot be built, it raises.
"""
import ctypes
import os

import numpy as np

from . import build as _build

c_u32 = ctypes.c_uint32
c_size = ctypes.c_size_t
c_p = ctypes.c_void_p
c_int = ctypes.c_int
c_f = ctypes.c_float
P = ctypes.POINTER

UVT_OK = 0
UVT_ERR_INVALID, UVT_ERR_CUDA, UVT_ERR_NO_DEVICE, UVT_ERR_OOM, UVT_ERR_FORMAT, UVT_ERR_IO = -1, -2, -3, -4, -5, -6
UVT_FLAG_HIT_BUFFER, UVT_FLAG_ENTITIES, UVT_FLAG_NO_DENSE, UVT_FLAG_FUSED_FRAME = 1, 2, 4, 8
UVT_ENTITY_BOXES, UVT_ENTITY_MODELS = 0, 1
UVT_LAYOUT_COMPACT, UVT_LAYOUT_REFERENCE = 0, 1
UVT_SCHED_TILE, UVT_SCHED_POOL = 0, 1
UVT_PIPELINE_PRIMARY, UVT_PIPELINE_SECONDARY, UVT_PIPELINE_EDIT, UVT_PIPELINE_BLIT = 0, 1, 2, 3
UVT_BUF_ALBEDO, UVT_BUF_NORMAL, UVT_BUF_POSITION, UVT_BUF_ILLUMINATION, UVT_BUF_FRAME, UVT_BUF_HIT = range(6)


class Params(ctypes.Structure):
    _fields_ = [("map_dim", c_u32), ("primary_max_steps", c_u32), ("shadow_max_steps", c_u32), ("edit_max_steps", c_u32),
                ("epsilon", c_f), ("flags", c_u32), ("layout", c_u32), ("scheduler", c_u32), ("reserved", c_u32 * 8)]


class Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ("rays", "t_in", "t_chunk", "t_block", "hits", "early_out")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class CameraState(ctypes.Structure):
    _fields_ = [("fov", c_f), ("pitch", c_f), ("yaw", c_f), ("cam_mat", c_f * 16), ("cam_pos", c_f * 4)]


CAMERA_DTYPE = np.dtype([("cam_pos", "<f4", 4), ("cam_mat", "<f4", 16), ("fov", "<f4"), ("_pad", "<f4", 3)])
HIT_DTYPE = np.dtype([("px", "<u4"), ("py", "<u4"), ("pz", "<u4"), ("block", "<u4"), ("color", "<u4"),
                      ("distance", "<f4"), ("trips", "<u2"), ("face", "u1"), ("exit_kind", "u1")])
assert CAMERA_DTYPE.itemsize == 96 and HIT_DTYPE.itemsize == 28

# name -> (restype, argtypes): every symbol include/*.h declares
SIGNATURES = {
    # uvt.h
    "uvt_default_params": (None, [P(Params)]),
    "uvt_create": (c_int, [P(Params), c_int, P(c_p)]),
    "uvt_destroy": (None, [c_p]),
    "uvt_last_error": (ctypes.c_char_p, [c_p]),
    "uvt_abi_version": (c_int, []),
    "uvt_set_stream": (c_int, [c_p, c_p]),
    "uvt_get_params": (c_int, [c_p, P(Params)]),
    "uvt_set_layout": (c_int, [c_p, c_u32]),
    "uvt_set_scheduler": (c_int, [c_p, c_u32]),
    "uvt_effective_layout": (c_int, [c_p]),
    "uvt_set_max_steps": (c_int, [c_p, c_u32, c_u32]),
    "uvt_pipeline_create": (c_int, [c_p, c_int, P(c_p)]),
    "uvt_pipeline_destroy": (None, [c_p]),
    "uvt_pipeline_dispatch": (c_int, [c_p, c_u32, c_u32, c_u32]),
    "uvt_world_alloc": (c_int, [c_p, c_u32, P(c_p), P(c_p), c_size]),
    "uvt_world_use_staging": (c_int, [c_p, c_u32, c_p, c_p, c_size]),
    "uvt_world_grow": (c_int, [c_p, c_size, P(c_p)]),
    "uvt_world_commit": (c_int, [c_p, c_size]),
    "uvt_world_commit_region": (c_int, [c_p, c_size, P(c_u32 * 3), P(c_u32 * 3)]),
    "uvt_world_set_voxel": (c_int, [c_p, c_u32, c_u32, c_u32, c_u32, P(c_int)]),
    "uvt_world_layout_checksum": (c_int, [c_p, P(ctypes.c_uint64 * 4)]),
    "uvt_atlas_upload": (c_int, [c_p, c_u32, c_u32, c_u32, c_u32, c_u32, c_u32, c_p]),
    "uvt_set_camera": (c_int, [c_p, c_p]),
    "uvt_set_cameras": (c_int, [c_p, c_p, c_int]),
    "uvt_resize": (c_int, [c_p, c_u32, c_u32]),
    "uvt_set_partition": (c_int, [c_p, c_u32, c_u32, c_u32]),
    "uvt_local_rows": (c_int, [c_p, P(c_u32)]),
    "uvt_dispatch_primary": (c_int, [c_p]),
    "uvt_dispatch_secondary": (c_int, [c_p]),
    "uvt_shade": (c_int, [c_p]),
    "uvt_dispatch_secondary_shade": (c_int, [c_p]),
    "uvt_dispatch_frame": (c_int, [c_p]),
    "uvt_set_entity_mode": (c_int, [c_p, c_u32]),
    "uvt_set_entities": (c_int, [c_p, c_p, c_u32]),
    "uvt_entity_model_upload": (c_int, [c_p, c_u32, c_p, c_u32]),
    "uvt_set_frame_chunks": (c_int, [c_p, c_u32]),
    "uvt_pick": (c_int, [c_p, c_p]),
    "uvt_sync": (c_int, [c_p]),
    "uvt_readback": (c_int, [c_p, c_int, c_p, c_size]),
    "uvt_buffer_bytes": (c_size, [c_p, c_int]),
    "uvt_readback_async": (c_int, [c_p, c_int, c_p, c_size]),
    "uvt_readback_wait": (c_int, [c_p]),
    "uvt_readback_bands_async": (c_int, [c_p, c_p, c_size]),
    "uvt_host_register": (c_int, [c_p, c_p, c_size]),
    "uvt_host_unregister": (c_int, [c_p, c_p]),
    "uvt_nccl_unique_id": (c_int, [c_p]),
    "uvt_nccl_init": (c_int, [c_p, c_p, c_int, c_int]),
    "uvt_nccl_shutdown": (c_int, [c_p]),
    "uvt_dispatch_frame_nccl": (c_int, [c_p, c_p, c_u32]),
    "uvt_device_ptr": (c_int, [c_p, c_int, P(c_p)]),
    "uvt_bind_frame_target": (c_int, [c_p, c_p, c_u32, c_u32]),
    "uvt_deinterleave": (c_int, [c_p, c_p, c_p, c_u32]),
    "uvt_shared_frame_create": (c_int, [c_p, P(c_p), c_p]),
    "uvt_shared_frame_open": (c_int, [c_p, c_p, P(c_p)]),
    "uvt_shared_frame_close": (c_int, [c_p, c_p]),
    "uvt_read_device": (c_int, [c_p, c_p, c_p, c_size]),
    "uvt_alloc_pinned": (c_int, [c_p, c_size, P(c_p)]),
    "uvt_free_pinned": (c_int, [c_p, c_p]),
    "uvt_count_pass": (c_int, [c_p, c_int, P(Counters)]),
    "uvt_last_pass_ms": (c_int, [c_p, c_int, P(c_f)]),
    "uvt_enable_timing": (c_int, [c_p, c_int]),
    "uvt_launch_count": (ctypes.c_uint64, [c_p]),
    "uvt_measure_l2_read_gbps": (c_int, [c_p, c_size, c_int, P(c_f)]),
    "uvt_measure_hbm_copy_gbps": (c_int, [c_p, c_size, c_int, P(c_f)]),
    "uvt_group_create": (c_int, [P(Params), P(c_int), c_int, P(c_p)]),
    "uvt_group_destroy": (None, [c_p]),
    "uvt_group_size": (c_int, [c_p]),
    "uvt_group_member": (c_p, [c_p, c_int]),
    "uvt_group_last_error": (ctypes.c_char_p, [c_p]),
    "uvt_group_world_alloc": (c_int, [c_p, c_u32, P(c_p), P(c_p), c_size]),
    "uvt_group_world_grow": (c_int, [c_p, c_size, P(c_p)]),
    "uvt_group_world_commit": (c_int, [c_p, c_size]),
    "uvt_group_world_commit_region": (c_int, [c_p, c_size, P(c_u32 * 3), P(c_u32 * 3)]),
    "uvt_group_atlas_upload": (c_int, [c_p, c_u32, c_u32, c_u32, c_u32, c_u32, c_u32, c_p]),
    "uvt_group_set_entity_mode": (c_int, [c_p, c_u32]),
    "uvt_group_set_entities": (c_int, [c_p, c_p, c_u32]),
    "uvt_group_entity_model_upload": (c_int, [c_p, c_u32, c_p, c_u32]),
    "uvt_group_set_camera": (c_int, [c_p, c_p]),
    "uvt_group_resize": (c_int, [c_p, c_u32, c_u32]),
    "uvt_group_dispatch_frame": (c_int, [c_p]),
    "uvt_group_sync": (c_int, [c_p]),
    "uvt_group_readback_frame": (c_int, [c_p, c_p, c_size]),
    "uvt_group_frame_ptr": (c_int, [c_p, P(c_p)]),
    "uvt_group_count_pass": (c_int, [c_p, c_int, P(Counters)]),
    # uvt_host.h
    "uvt_brickmap_create": (c_int, [c_p, c_u32, P(c_p)]),
    "uvt_brickmap_create_group": (c_int, [c_p, c_u32, P(c_p)]),
    "uvt_brickmap_destroy": (None, [c_p]),
    "uvt_brickmap_clear": (None, [c_p]),
    "uvt_brickmap_set": (c_int, [c_p, c_u32, c_u32, c_u32, c_u32]),
    "uvt_brickmap_get": (c_u32, [c_p, c_u32, c_u32, c_u32]),
    "uvt_brickmap_is_walkable": (c_int, [c_p, c_u32, c_u32, c_u32]),
    "uvt_brickmap_dim": (c_u32, [c_p]),
    "uvt_brickmap_n_bricks": (c_size, [c_p]),
    "uvt_brickmap_capacity": (c_size, [c_p]),
    "uvt_brickmap_chunks": (c_p, [c_p]),
    "uvt_brickmap_bricks": (c_p, [c_p]),
    "uvt_brickmap_bind": (c_int, [c_p]),
    "uvt_brickmap_mark_dirty": (None, [c_p]),
    "uvt_brickmap_save": (c_int, [c_p, ctypes.c_char_p]),
    "uvt_brickmap_load": (c_int, [c_p, ctypes.c_char_p, P(c_p)]),
    "uvt_procgen": (c_int, [c_p, c_u32, c_f, c_f]),
    "uvt_procgen_device": (c_int, [c_p, c_u32, c_f, c_f]),
    "uvt_world_procgen_plan": (c_int, [c_p, c_f, c_f, P(c_size)]),
    "uvt_world_procgen_fill": (c_int, [c_p]),
    "uvt_noise2_fbm": (c_f, [c_f, c_f]),
    "uvt_procgen_height": (c_u32, [c_u32, c_u32, c_u32, c_f, c_f]),
    "uvt_vox_parse": (c_int, [c_p, c_size, P(c_p)]),
    "uvt_vox_open": (c_int, [ctypes.c_char_p, P(c_p)]),
    "uvt_vox_free": (None, [c_p]),
    "uvt_vox_n_models": (c_u32, [c_p]),
    "uvt_vox_model_size": (c_int, [c_p, c_u32, P(c_u32 * 3)]),
    "uvt_vox_model_n_voxels": (c_u32, [c_p, c_u32]),
    "uvt_vox_model_voxels": (c_p, [c_p, c_u32]),
    "uvt_vox_palette": (c_p, [c_p]),
    "uvt_vox_error": (ctypes.c_char_p, []),
    "uvt_atlas_create": (c_int, [c_p, P(c_p)]),
    "uvt_atlas_create_group": (c_int, [c_p, P(c_p)]),
    "uvt_atlas_destroy": (None, [c_p]),
    "uvt_atlas_load_block_model": (c_int, [c_p, ctypes.c_char_p]),
    "uvt_atlas_load_block_model_mem": (c_int, [c_p, c_p, c_size]),
    "uvt_atlas_append_model": (c_int, [c_p, c_p]),
    "uvt_atlas_current_index": (c_u32, [c_p]),
    "uvt_atlas_get_model": (c_int, [c_p, c_u32, c_p]),
    "uvt_camera_init": (None, [P(CameraState)]),
    "uvt_camera_rotate": (None, [P(CameraState), c_f, c_f]),
    "uvt_camera_set_pos": (None, [P(CameraState), P(c_f * 4)]),
    "uvt_camera_increment_fov": (None, [P(CameraState), c_f]),
    "uvt_camera_as_uniform_data": (None, [P(CameraState), c_p]),
    "uvt_mat_from_pitch_yaw": (None, [c_f, c_f, P(c_f * 16)]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load libuvt.so, building it first if it is missing or stale.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UVT_LIB_PATH") or _build.LIB  # UVT_LIB_PATH: an experimental variant built by build.build(out=...)
    if path == _build.LIB and (not os.path.exists(path) or (_build.is_stale() and os.environ.get("UVT_NO_REBUILD") != "1")):
        _build.build()
    L = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
