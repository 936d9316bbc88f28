"""unnamed-voxel-tracer_b200 — B200-native voxel ray-traversal pass behind the reference's
renderer entry points.  The compute path is hand-written sm_100a CUDA in csrc/ behind the C
ABI of include/uvt.h; this package is the Python mirror of the reference's host interface
(gfx / voxel / procgen / game) over that ABI.  There is no CPU fallback: without the built
library and a B200 every dispatch raises.
"""
from . import _native, build, gfx, voxel, procgen, game, tiles, scenes  # noqa: F401
from .gfx import Context, Group, UvtError, init  # noqa: F401

__all__ = ["gfx", "voxel", "procgen", "game", "tiles", "scenes", "Context", "Group", "UvtError", "init", "build"]
