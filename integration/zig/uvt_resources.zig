//! Camera uniform block, G-buffer, brickmap and model atlas over the C ABI.
const std = @import("std");
const uvt = @import("uvt.zig");
const c = uvt.c;

/// The camera UBO of the game (`PersistentMappedBuffer(Camera.UniformData)`): the game writes the 96-byte
/// block through deref() every frame and bind(8) publishes it.
pub fn CameraUniforms(comptime UniformData: type) type {
    comptime std.debug.assert(@sizeOf(UniformData) == @sizeOf(c.uvt_camera));
    return struct {
        data: UniformData = undefined,

        pub fn deref(self: *@This()) *UniformData {
            return &self.data;
        }

        pub fn bind(self: *@This(), _: u32) void {
            uvt.check(c.uvt_set_camera(uvt.ctx, @ptrCast(&self.data))) catch {};
        }
    };
}

pub const GBuffer = struct {
    width: u32,
    height: u32,

    pub fn init(width: u32, height: u32) @This() {
        uvt.check(c.uvt_resize(uvt.ctx, width, height)) catch {};
        return .{ .width = width, .height = height };
    }

    pub fn resize(self: *@This(), width: u32, height: u32) void {
        uvt.check(c.uvt_resize(uvt.ctx, width, height)) catch {};
        self.width = width;
        self.height = height;
    }

    pub fn bind_images(_: *@This(), _: u32) void {}
    pub fn bind_textures(_: *@This(), _: u32) void {}
    pub fn deinit(_: *@This()) void {}

    /// Presentation: copy the shaded frame into `dst` (RGBA8, row 0 = bottom row) for a PBO / texture upload.
    pub fn read_frame(self: *@This(), dst: []u8) void {
        std.debug.assert(dst.len >= @as(usize, self.width) * self.height * 4);
        uvt.check(c.uvt_readback(uvt.ctx, c.UVT_BUF_FRAME, dst.ptr, @as(usize, self.width) * self.height * 4)) catch {};
    }
};

pub fn VoxelBrickmap(comptime dim: comptime_int, comptime chsize: comptime_int) type {
    comptime std.debug.assert(chsize == 8);
    return struct {
        handle: ?*c.uvt_brickmap,
        dirty: bool = true,

        pub fn init() @This() {
            var h: ?*c.uvt_brickmap = null;
            uvt.check(c.uvt_brickmap_create(uvt.ctx, dim, &h)) catch {};
            return .{ .handle = h };
        }

        pub fn clear(self: *@This(), _: u32) void {
            c.uvt_brickmap_clear(self.handle);
            self.dirty = true;
        }

        pub fn set(self: *@This(), x: usize, y: usize, z: usize, voxel: u32) void {
            _ = c.uvt_brickmap_set(self.handle, @intCast(x), @intCast(y), @intCast(z), voxel);
            self.dirty = true;
        }

        pub fn get(self: *@This(), x: usize, y: usize, z: usize) u32 {
            return c.uvt_brickmap_get(self.handle, @intCast(x), @intCast(y), @intCast(z));
        }

        pub fn is_walkable(self: *@This(), x: usize, y: usize, z: usize) bool {
            return c.uvt_brickmap_is_walkable(self.handle, @intCast(x), @intCast(y), @intCast(z)) != 0;
        }

        /// GL mappings are live; CUDA staging is published here.  The library tracks the box of blocks written through
        /// set() since the last bind and publishes only that (uvt_world_commit_region, a fraction of a millisecond);
        /// the flag below merely skips the call on clean frames.
        pub fn bind(self: *@This(), _: u32) void {
            if (!self.dirty) return;
            uvt.check(c.uvt_brickmap_bind(self.handle)) catch {};
            self.dirty = false;
        }
    };
}

pub const VoxelModelAtlas = struct {
    handle: ?*c.uvt_atlas,

    pub fn init() @This() {
        var h: ?*c.uvt_atlas = null;
        uvt.check(c.uvt_atlas_create(uvt.ctx, &h)) catch {};
        return .{ .handle = h };
    }

    pub fn load_block_model(self: *@This(), model: [:0]const u8, _: std.mem.Allocator) !void {
        if (c.uvt_atlas_load_block_model(self.handle, model.ptr) != c.UVT_OK) {
            std.log.err("uvt: {s}: {s}", .{ model, c.uvt_vox_error() });
            return error.InvalidVoxFile;
        }
    }

    pub fn bind(_: *@This(), _: u32) void {}
};

/// procgen.zig's entry point.  A fresh map bound to a ctx is generated on the GPU (same LCG stream, byte-identical world:
/// uvt_procgen_device); anything the device path does not cover falls back to the serial host version on the pinned staging.
pub fn procgen(comptime dim: comptime_int, world: anytype, offsetX: f32, offsetY: f32) void {
    if (c.uvt_procgen_device(world.handle, dim, offsetX, offsetY) != c.UVT_OK)
        _ = c.uvt_procgen(world.handle, dim, offsetX, offsetY);
    world.dirty = true;
}

/// traceEntities (map.glsl:172-248).  Nothing to call for the reference as it runs (five literal boxes, shadow pass only).
/// `enableEntityModels` switches on what the reference keeps dead / commented: the sub-model DDA and the primary-pass
/// composite, with the entity model the commented `models.load_model("assets/chicken.vox", allocator, 32)` of game.zig:114
/// would have loaded (texels: size^3 RGBA8, x fastest, then y, then z).
pub fn enableEntityModels(positions: []const [3]f32, texels: ?[]const u32, size: u32) void {
    uvt.check(c.uvt_set_entities(uvt.ctx, @ptrCast(positions.ptr), @intCast(positions.len))) catch {};
    if (texels) |t| uvt.check(c.uvt_entity_model_upload(uvt.ctx, size, t.ptr, 8 * size)) catch {};
    uvt.check(c.uvt_set_entity_mode(uvt.ctx, c.UVT_ENTITY_MODELS)) catch {};
}
