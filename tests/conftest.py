import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ASSETS = "/root/reference/assets"  # exists only in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def uvt():
    """The product package (its directory name carries a hyphen, hence importlib)."""
    return importlib.import_module("unnamed-voxel-tracer_b200")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def models():
    return np.load(os.path.join(GOLDEN, "atlas_models.npy"))


@pytest.fixture(scope="session")
def atlas(oracle, models):
    return oracle.atlas_from_models(models)


class Scene:
    """A host-built world + atlas for the oracle; with a ctx the brickmap storage IS the ctx's
    pinned staging and the scene is published through the C ABI (commit + atlas uploads)."""

    def __init__(self, uvt, oracle, dim, models, atlas, fill=None, ctx=None):
        self.uvt, self.dim, self.models, self.ctx = uvt, dim, models, ctx
        self.bm = uvt.voxel.VoxelBrickmap.init(dim, 8, ctx)
        if fill == "procgen":
            uvt.procgen.procgen(dim, self.bm)
        elif callable(fill):
            fill(self.bm)
        self.chunks = self.bm.chunks().copy()
        self.bricks = self.bm.bricks().copy()
        self.n_bricks = self.bm.n_bricks
        self.oracle_world = oracle.World(dim, self.chunks, self.bricks, atlas)
        if ctx is not None:
            self.atlas = uvt.voxel.VoxelModelAtlas.init(ctx)
            for m in models:
                self.atlas.append_model(m)
            self.bm.bind(9)


@pytest.fixture(scope="session")
def scene_factory(uvt, oracle, models, atlas):
    cache = {}

    def make(dim, fill=None, key=None, ctx=None):
        if ctx is not None:
            return Scene(uvt, oracle, dim, models, atlas, fill, ctx)
        k = (dim, key if key is not None else (fill if isinstance(fill, str) else id(fill)))
        if k not in cache:
            cache[k] = Scene(uvt, oracle, dim, models, atlas, fill)
        return cache[k]

    return make


@pytest.fixture(scope="session")
def world512(scene_factory):
    """W1: procgen(512) with the reference seeds."""
    return scene_factory(512, "procgen")


@pytest.fixture(scope="session")
def world64(scene_factory):
    return scene_factory(64, "procgen")


def pitch_yaw_matrix(uvt, pitch, yaw):
    import ctypes
    m = (ctypes.c_float * 16)()
    uvt._native.load().uvt_mat_from_pitch_yaw(float(pitch), float(yaw), ctypes.byref(m))
    return np.array(m, dtype=np.float32)


def camera_k0(oracle):
    """K0: spawn (256,22,256) + (0,3,0), identity, fov pi/2 (game.zig:40,212; camera.zig:6-8)."""
    return oracle.make_camera((256.0, 25.0, 256.0))


def camera_k1(uvt, oracle):
    """K1: a pitched/yawed view over the hills so trees and shadows are exercised."""
    return oracle.make_camera((200.0, 48.0, 140.0), pitch_yaw_matrix(uvt, 0.35, 0.6))
