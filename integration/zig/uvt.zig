//! Root of the CUDA-backed `gfx` replacement: the C ABI of include/uvt.h seen from Zig.
//! NOT compiled in the build image (no Zig toolchain there); the same call order is exercised
//! from C by tests/c/game_loop.c and from Python by unnamed-voxel-tracer_b200/game.py.
const std = @import("std");

pub const c = @cImport({
    @cInclude("uvt.h");
    @cInclude("uvt_host.h");
});

pub var ctx: ?*c.uvt_ctx = null;

pub const Error = error{ GraphicsInitFailed, ShaderCompilationError, DeviceError };

/// Stands in for gfx.init(window) + enableDebug(): one CUDA device, one in-order stream.
pub fn init(device: c_int) Error!void {
    var params: c.uvt_params = undefined;
    c.uvt_default_params(&params);
    if (c.uvt_create(&params, device, &ctx) != c.UVT_OK) {
        std.log.err("uvt: {s}", .{c.uvt_last_error(null)});
        return Error.GraphicsInitFailed;
    }
    std.log.info("uvt: CUDA renderer ready (ABI {})", .{c.uvt_abi_version()});
}

pub fn deinit() void {
    c.uvt_destroy(ctx);
    ctx = null;
}

/// Per-frame calls return void in the reference; a failing C call is logged and surfaced as DeviceError.
pub fn check(rc: c_int) Error!void {
    if (rc != c.UVT_OK) {
        std.log.warn("uvt: {s}", .{c.uvt_last_error(ctx)});
        return Error.DeviceError;
    }
}

/// gfx.resize(width, height): the viewport follows the G-buffer, nothing to do.
pub fn resize(_: u32, _: u32) void {}

/// gfx.clear(r, g, b): the blit writes every pixel of the frame.
pub fn clear(_: f32, _: f32, _: f32) void {}
