//! ComputePipeline / RasterPipeline with the public shape the game code uses
//! (init, bind, dispatch / draw, deinit), backed by precompiled sm_100a kernels.
const std = @import("std");
const uvt = @import("uvt.zig");
const c = uvt.c;

fn kindOf(path: []const u8) ?c_uint {
    if (std.mem.endsWith(u8, path, "primary.comp.glsl")) return c.UVT_PIPELINE_PRIMARY;
    if (std.mem.endsWith(u8, path, "secondary.comp.glsl")) return c.UVT_PIPELINE_SECONDARY;
    if (std.mem.endsWith(u8, path, "terrain_edit.comp.glsl")) return c.UVT_PIPELINE_EDIT;
    return null;
}

pub const ComputePipeline = struct {
    handle: ?*c.uvt_pipeline,

    /// `file` names the reference shader whose pass this pipeline runs; nothing is compiled at run time.
    pub fn init(_: std.mem.Allocator, file: []const u8) uvt.Error!@This() {
        const kind = kindOf(file) orelse return uvt.Error.ShaderCompilationError;
        var h: ?*c.uvt_pipeline = null;
        if (c.uvt_pipeline_create(uvt.ctx, kind, &h) != c.UVT_OK) return uvt.Error.ShaderCompilationError;
        return .{ .handle = h };
    }

    pub fn bind(_: *const @This()) void {}

    /// Stream-ordered like glDispatchCompute + glMemoryBarrier: the next pass sees this one's images.
    pub fn dispatch(self: *const @This(), x: c_uint, y: c_uint, z: c_uint) void {
        uvt.check(c.uvt_pipeline_dispatch(self.handle, x, y, z)) catch {};
    }

    pub fn deinit(self: *const @This()) void {
        c.uvt_pipeline_destroy(self.handle);
    }
};

pub const RasterPipeline = struct {
    handle: ?*c.uvt_pipeline,

    pub fn init(_: std.mem.Allocator, _: []const u8, _: []const u8) uvt.Error!@This() {
        var h: ?*c.uvt_pipeline = null;
        if (c.uvt_pipeline_create(uvt.ctx, c.UVT_PIPELINE_BLIT, &h) != c.UVT_OK) return uvt.Error.ShaderCompilationError;
        return .{ .handle = h };
    }

    pub fn bind(_: *const @This()) void {}

    /// The full-screen strip: shades the G-buffer into the frame (crosshair and vignette included).
    pub fn draw(self: *@This(), _: usize) void {
        uvt.check(c.uvt_pipeline_dispatch(self.handle, 1, 1, 1)) catch {};
    }

    pub fn deinit(self: *const @This()) void {
        c.uvt_pipeline_destroy(self.handle);
    }
};
