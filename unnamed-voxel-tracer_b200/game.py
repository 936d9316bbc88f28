"""Mirror of the renderer-facing part of src/game.zig: init order, pre_render, render.

Physics, audio and input (game.zig:131-221) are out of scope; what is kept is exactly the
sequence of gfx calls the reference issues, so this class is also the "C test that exercises
the same call order" the Zig glue of SURVEY §8f-1 would replace.
"""
import os

import numpy as np

from . import gfx, procgen, voxel

# load order of src/game.zig:101-113
BLOCK_MODEL_FILES = ["models.vox", "grass.vox", "grass2.vox", "grass3.vox", "grass4.vox", "grass5.vox", "rock.vox",
                     "flower.vox", "water.vox", "tree.vox", "leaves.vox", "dirt.vox", "sand.vox"]


class Game:
    def __init__(self, ctx, *, dim=512, width=1280, height=720, assets_dir=None, models=None, world=None):
        """game_init (game.zig:54-129).  `models` ([n,512] texels) replaces reading .vox files;
        `world` (a committed-able VoxelBrickmap) replaces procgen."""
        self.ctx = ctx
        self.primary_trace_pipeline = gfx.ComputePipeline.init(ctx, "assets/shaders/primary.comp.glsl")
        self.secondary_trace_pipeline = gfx.ComputePipeline.init(ctx, "assets/shaders/secondary.comp.glsl")
        self.edit_pipeline = gfx.ComputePipeline.init(ctx, "assets/shaders/terrain_edit.comp.glsl")
        self.raster_pipeline = gfx.RasterPipeline.init(ctx, "assets/shaders/blit.vertex.glsl", "assets/shaders/blit.fragment.glsl")
        self.gbuffer = gfx.GBuffer.init(ctx, width, height)
        self.cam_uniforms = gfx.PersistentMappedBuffer.init(ctx)
        if world is None:
            world = voxel.VoxelBrickmap.init(dim, 8, ctx)
            procgen.procgen(dim, world, 0.0, 0.0)
        self.voxels = world
        self.models = voxel.VoxelModelAtlas.init(ctx)
        if models is not None:
            for m in np.asarray(models, dtype=np.uint32).reshape(-1, 512):
                self.models.append_model(m)
        else:
            for f in BLOCK_MODEL_FILES:
                self.models.load_block_model(os.path.join(assets_dir, f))
        self.cam = gfx.Camera()
        self.position = np.array([256.0, 22.0, 256.0, 0.0], dtype=np.float32)  # game.zig:40

    def update(self):
        """game.zig:207-212 (camera part)."""
        self.cam.set_pos(self.position + np.array([0.0, 3.0, 0.0, 0.0], dtype=np.float32))

    def pre_render(self):
        """game.zig:224-229."""
        self.cam_uniforms.deref()[...] = self.cam.as_uniform_data()

    def render(self):
        """game.zig:232-256, call for call."""
        self.cam_uniforms.bind(8)
        self.voxels.bind(9)
        self.models.bind(6)
        self.gbuffer.bind_images(0)
        workgroup_size_x = self.gbuffer.width // 32 + 1
        workgroup_size_y = self.gbuffer.height // 32 + 1
        self.primary_trace_pipeline.bind()
        self.primary_trace_pipeline.dispatch(workgroup_size_x, workgroup_size_y, 1)
        self.secondary_trace_pipeline.bind()
        self.secondary_trace_pipeline.dispatch(workgroup_size_x, workgroup_size_y, 1)
        self.gbuffer.bind_textures(0)
        self.raster_pipeline.bind()
        self.raster_pipeline.draw(4)

    def window_resized(self, width, height):
        """game.zig:197-205."""
        self.gbuffer.resize(width, height)

    def deinit(self):
        for p in (self.primary_trace_pipeline, self.secondary_trace_pipeline, self.edit_pipeline, self.raster_pipeline):
            p.deinit()
