// kernels.cuh — the three passes of the reference frame (src/game.zig:244-255) as sm_100a
// kernels: primary (primary.comp.glsl), secondary (secondary.comp.glsl), shade
// (blit.fragment.glsl), plus the fused frame kernel and small utilities.
#pragma once

#include "trace.cuh"

namespace uvt {

// camera.glsl:8
#define UVT_SUN_X 7.52185881e-01f
#define UVT_SUN_Y 6.58950984e-01f
#define UVT_SUN_Z 7.52185881e-01f

// Camera as the kernels consume it: camera.glsl:2-6 with tan(fov/2) evaluated once on the host.
struct CamDev {
    float pos[3];
    float tan_half_fov;
    float mat[16];  // row j = GLSL column j
};

// Image geometry + multi-GPU row-band partition (uvt_set_partition).
struct ViewDev {
    uint32_t W, H;          // full frame
    uint32_t local_rows;    // rows stored by this ctx
    uint32_t row0;          // first local row of this launch (band groups of uvt_dispatch_frame_nccl; 0 otherwise)
    uint32_t band_rows, n_parts, part;
    uint32_t map_dim;
    uint32_t max_steps;
    float epsilon;
    uint32_t entities;
    // local row -> global row; returns false for padding rows
    __device__ __forceinline__ bool global_row(uint32_t ly, uint32_t &y) const {
        if (n_parts == 1u) { y = ly; return ly < H; }
        const uint32_t lb = ly / band_rows;
        y = (lb * n_parts + part) * band_rows + (ly - lb * band_rows);
        return ly < local_rows && y < H;
    }
};

struct GBufDev {
    uint32_t *albedo;   // RGBA8
    uint32_t *normal;   // RGBA8
    float4 *position;   // RGBA32F
    uint32_t *illum;    // RGBA8
    uint32_t *frame;    // RGBA8
    uint8_t *hit;       // uvt_hit[...] (28 B each) or nullptr
    size_t layer_pixels;  // pixels per layer (W * local_rows)
};

// RGBA8 UNORM store conversion: clamp, *255, +0.5, truncate (0.3 -> 77)
__device__ __forceinline__ uint32_t unorm8(float c) {
    if (c != c) return 0u;
    c = gmin(gmax(c, 0.0f), 1.0f);
    return __float2uint_rz(c * 255.0f + 0.5f);
}
__device__ __forceinline__ uint32_t pack_rgba8(float r, float g, float b, float a) {
    return unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (unorm8(a) << 24);
}

// RGBA8 UNORM fetch: c / 255 (correctly rounded).  The normals the primary pass stores are 0 or 255: no division for those.
__device__ __forceinline__ float unorm8_to_float(uint32_t c) {
    if (c == 0u) return 0.0f;
    if (c == 255u) return 1.0f;
    return (float)c / 255.0f;
}

// SkyDome2: camera.glsl:11-19 (rgb; alpha is 1).  pow(sun, 8) and pow(sun, 3) are integer
// powers evaluated by multiplication (within the 1/255 shading tolerance of the oracle's powf).
__device__ __forceinline__ void sky_dome2(float rx, float ry, float rz, float &r, float &g, float &b) {
    const float sl = sqrtf(UVT_SUN_X * UVT_SUN_X + UVT_SUN_Y * UVT_SUN_Y + UVT_SUN_Z * UVT_SUN_Z);
    const float rl = sqrtf(rx * rx + ry * ry + rz * rz);
    const float dot = (UVT_SUN_X / sl) * (rx / rl) + (UVT_SUN_Y / sl) * (ry / rl) + (UVT_SUN_Z / sl) * (rz / rl);
    const float sun = gmin(gmax(dot, 0.0f), 1.2f);
    const float s2 = sun * sun, s4 = s2 * s2, p8 = s4 * s4, p3 = s2 * sun;
    const float k = ry * 0.2f;
    r = 0.6f - k * 1.0f + 0.15f * 0.5f;
    g = 0.71f - k * 0.5f + 0.15f * 0.5f;
    b = 0.75f - k * 1.0f + 0.15f * 0.5f;
    r += 0.4f * 1.0f * p8; g += 0.4f * 0.6f * p8; b += 0.4f * 0.1f * p8;
    r += 0.2f * p3; g += 0.08f * p3; b += 0.04f * p3;
}

// intersectAABB: map.glsl:21-29
__device__ __forceinline__ void intersect_aabb(float ox, float oy, float oz, float dx, float dy, float dz,
                                               float lx, float ly, float lz, float hx, float hy, float hz,
                                               float &t_near, float &t_far) {
    const float ax = (lx - ox) / dx, bx = (hx - ox) / dx;
    const float ay = (ly - oy) / dy, by = (hy - oy) / dy;
    const float az = (lz - oz) / dz, bz = (hz - oz) / dz;
    t_near = gmax(gmax(gmin(ax, bx), gmin(ay, by)), gmin(az, bz));
    t_far = gmin(gmin(gmax(ax, bx), gmax(ay, by)), gmax(az, bz));
}

// traceEntities live part: map.glsl:172-201
__device__ __forceinline__ bool trace_entities(float ox, float oy, float oz, float dx, float dy, float dz, float max_distance) {
    const float P[5][3] = {{256.f, 21.f, 256.f}, {251.f, 21.f, 259.f}, {253.f, 21.f, 256.f}, {251.f, 21.f, 256.f}, {257.f, 21.f, 261.f}};
    float prev_d = __int_as_float(0x7f800000);
    bool any = false;
    const float dd = dx * dx + dy * dy + dz * dz;
    {   // the five box centres lie within 3.91 blocks of (254.5, 21.5, 259): a line that passes that point by more than 5 blocks is
        // more than 1.09 from every centre, so the per-box rejection below (distance > 1) would drop all five — which is
        // every shadow ray of a frame that looks somewhere else
        const float vx = 254.5f - ox, vy = 21.5f - oy, vz = 259.0f - oz;
        const float cx = vy * dz - vz * dy, cy = vz * dx - vx * dz, cz = vx * dy - vy * dx;
        if (cx * cx + cy * cy + cz * cz > 25.0f * dd) return false;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        {   // A line further than 1 block from the box centre misses the unit box (half diagonal 0.866): its slab intervals are
            // apart by more than 0.13 / |d| in parameter, far beyond the rounding of the exact test below (< 1e-3 at map scale),
            // so that test would find tf < tn.  The cross product decides it without the six divisions.
            const float vx = P[i][0] + 0.5f - ox, vy = P[i][1] + 0.5f - oy, vz = P[i][2] + 0.5f - oz;
            const float cx = vy * dz - vz * dy, cy = vz * dx - vx * dz, cz = vx * dy - vy * dx;
            if (cx * cx + cy * cy + cz * cz > dd) continue;
        }
        const float ex = ox - P[i][0], ey = oy - P[i][1], ez = oz - P[i][2];
        const float dist = sqrtf(ex * ex + ey * ey + ez * ez);
        if (dist >= max_distance) continue;
        float tn, tf;
        intersect_aabb(ox, oy, oz, dx, dy, dz, P[i][0], P[i][1], P[i][2], P[i][0] + 1.0f, P[i][1] + 1.0f, P[i][2] + 1.0f, tn, tf);
        if (tf >= tn && prev_d >= tf) {  // the re-test at :197-199 repeats this box's own tf >= tn
            any = true;
            prev_d = tf;
        }
    }
    return any;
}

// primary.comp.glsl:31-43: camera ray and the traceMap start point
__device__ __forceinline__ void primary_ray(const CamDev &cam, const ViewDev &v, uint32_t px, uint32_t py,
                                            float &dx, float &dy, float &dz, float &sx, float &sy, float &sz) {
    float ux = (float)px / (float)v.W * 2.0f - 1.0f;
    float uy = (float)py / (float)v.H * 2.0f - 1.0f;
    uy *= (float)v.H / (float)v.W;
    ux *= cam.tan_half_fov;
    uy *= cam.tan_half_fov;
    float q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float acc = cam.mat[0 * 4 + i] * ux;
        acc = acc + cam.mat[1 * 4 + i] * uy;
        acc = acc + cam.mat[2 * 4 + i] * 1.0f;
        acc = acc + cam.mat[3 * 4 + i] * 1.0f;
        q[i] = acc;
    }
    const float len = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    dx = q[0] / len; dy = q[1] / len; dz = q[2] / len;
    const float D = (float)v.map_dim;
    // A camera strictly inside the map (the game's case) has tNear < 0 for every finite direction: per axis one of (0 - o) / d,
    // (D - o) / d is negative (or -inf for d = +-0), so max(tNear, 0) = 0 and the start point is o + d * 0 - EPSILON = o - EPSILON
    // (o != 0; d * 0 = +-0).  The six divisions of the slab test are only run when that does not hold.
    const bool inside = cam.pos[0] > 0.0f && cam.pos[0] < D && cam.pos[1] > 0.0f && cam.pos[1] < D && cam.pos[2] > 0.0f && cam.pos[2] < D;
    const float inf = __int_as_float(0x7f800000);
    if (inside && fabsf(dx) < inf && fabsf(dy) < inf && fabsf(dz) < inf) {
        sx = cam.pos[0] - v.epsilon;
        sy = cam.pos[1] - v.epsilon;
        sz = cam.pos[2] - v.epsilon;
        return;
    }
    float tn, tf;
    intersect_aabb(cam.pos[0], cam.pos[1], cam.pos[2], dx, dy, dz, 0.0f, 0.0f, 0.0f, D, D, D, tn, tf);
    const float t0 = gmax(tn, 0.0f);
    sx = cam.pos[0] + dx * t0 - v.epsilon;
    sy = cam.pos[1] + dy * t0 - v.epsilon;
    sz = cam.pos[2] + dz * t0 - v.epsilon;
}

// packed uvt_hit record (28 B): 7 words
__device__ __forceinline__ void store_hit(uint8_t *hit_base, size_t i, const Hit &h, float distance) {
    uint32_t *p = reinterpret_cast<uint32_t *>(hit_base + i * 28u);
    p[0] = h.px; p[1] = h.py; p[2] = h.pz;
    p[3] = h.block;
    p[4] = h.data;
    p[5] = __float_as_uint(distance);
    p[6] = (h.trips & 0xFFFFu) | ((h.face & 0xFFu) << 16) | ((h.exit_kind & 0xFFu) << 24);
}

__device__ __forceinline__ uint32_t normal_rgba8(uint32_t face) {
    // normals[face-1] stored to RGBA8 UNORM: negative components clamp to 0, alpha = 255
    // faces 2,4,6 are +x,+y,+z; faces 1,3,5 store (0,0,0,255)
    uint32_t n = 0xFF000000u;
    if (face == 2u) n |= 0x000000FFu;
    if (face == 4u) n |= 0x0000FF00u;
    if (face == 6u) n |= 0x00FF0000u;
    return n;
}

struct DevCounters {
    unsigned long long rays, t_in, t_chunk, t_block, hits, early_out;
};

__device__ __forceinline__ void warp_add(unsigned long long *dst, uint32_t v) {
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31u) == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

// CTA tile: 128 threads = 4 warps in a 2x2 arrangement; a warp covers kWarpW x kWarpH pixels (8x4 by default:
// the most coherent footprint, SIMT efficiency 97.8 % at trip granularity), the CTA 2*kWarpW x 2*kWarpH.
#ifndef UVT_WARP_W
#define UVT_WARP_W 8
#endif
#ifndef UVT_MIN_BLOCKS
#define UVT_MIN_BLOCKS 9  // <= 56 registers, 36 warps/SM: measured best of {7, 8, 9, 10} on the 1080p primary pass (0.353 ms vs 0.374-0.390)
#endif
#ifndef UVT_MIN_BLOCKS_FRAME
#define UVT_MIN_BLOCKS_FRAME 7  // the fused frame kernel keeps two rays' worth of state: <= 73 registers measured faster than <= 51
#endif
constexpr int kWarpW = UVT_WARP_W, kWarpH = 32 / UVT_WARP_W;
#ifndef UVT_CTA_WARPS
#define UVT_CTA_WARPS 4  // warps per CTA of the pixel-per-thread kernels, two abreast (experiment knob: 2 = 64-thread CTAs)
#endif
constexpr int kThreads = 128;                     // pooled-scheduler kernels
constexpr int kTileThreads = 32 * UVT_CTA_WARPS;  // pixel-per-thread kernels
constexpr int kTileW = 2 * kWarpW, kTileH = (UVT_CTA_WARPS / 2) * kWarpH;
static_assert(UVT_CTA_WARPS == 2 || UVT_CTA_WARPS == 4 || UVT_CTA_WARPS == 8, "warps are laid out two abreast");
static_assert(kWarpW * kWarpH == 32, "a warp covers 32 pixels");

__device__ __forceinline__ void tile_pixel(uint32_t &x, uint32_t &ly) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    x = blockIdx.x * kTileW + (warp & 1u) * kWarpW + (lane % kWarpW);
    ly = blockIdx.y * kTileH + (warp >> 1) * kWarpH + (lane / kWarpW);
}

__device__ __forceinline__ void stage_masks(uint32_t *smem, const uint32_t *__restrict__ gmasks, uint32_t n_mats) {
#if !UVT_SMEM_MASKS
    return;
#endif
    // only the materials in use, at most kSmemMaskMats: (n_mats + 1) x 16 words (id 0 is the empty block)
    const uint32_t n = min(n_mats + 1u, (uint32_t)kSmemMaskMats) * 16u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) smem[i] = __ldg(&gmasks[i]);
    __syncthreads();
}

template <class World>
struct WorldArgs;

template <>
struct WorldArgs<WorldRef> {
    WorldRef w;
    const uint32_t *masks;  // unused
    uint32_t n_mats;
};
template <>
struct WorldArgs<WorldCompact> {
    WorldCompact w;
    const uint32_t *masks;  // global [256][16]
    uint32_t n_mats;        // material ids in use: 1..n_mats
};
template <>
struct WorldArgs<WorldDense> {
    WorldDense w;
    const uint32_t *masks;
    uint32_t n_mats;
};

// ---- primary pass ------------------------------------------------------------------------
template <class World, int COUNT, bool HITBUF, bool BATCH>
__global__ void __launch_bounds__(kTileThreads, UVT_MIN_BLOCKS) primary_kernel(WorldArgs<World> wa, const CamDev *__restrict__ cams, CamDev cam0,
                                                           ViewDev v, GBufDev gb, DevCounters *counters) {
    __shared__ uint32_t s_masks[(UVT_SMEM_MASKS && kIsCompact<World>) ? kSmemMaskMats * 16 : 1];
    World w = wa.w;
    if constexpr (kIsCompact<World>) {
        stage_masks(s_masks, wa.masks, wa.n_mats);
        w.smem_masks = s_masks;
    }
    uint32_t x, ly, y;
    tile_pixel(x, ly);
    ly += v.row0;
    const bool valid = v.global_row(ly, y) && x < v.W;
    const CamDev &cam = BATCH ? cams[blockIdx.z] : cam0;  // single camera: straight from the parameter bank

    TripCounts tc = {0, 0, 0};
    uint32_t is_hit = 0;
    float dx = 0.0f, dy = 0.0f, dz = 1.0f, sx = 0.0f, sy = 0.0f, sz = 0.0f;
    if (valid) primary_ray(cam, v, x, y, dx, dy, dz, sx, sy, sz);
    // traceMap's zero patch (map.glsl:85-90) applied here, remembering the patched components: the sky colour of a
    // miss needs the unpatched direction, and one mask register is cheaper to keep alive than three floats
    const uint32_t zmask = (dx == 0.0f ? 1u : 0u) | (dy == 0.0f ? 2u : 0u) | (dz == 0.0f ? 4u : 0u);
    if (zmask & 1u) dx = 0.001f;
    if (zmask & 2u) dy = 0.001f;
    if (zmask & 4u) dz = 0.001f;
    Hit h;
    trace<World, COUNT>(w, valid, sx, sy, sz, dx, dy, dz, (int)v.max_steps, (int)(8u * v.map_dim), h, tc);  // all 32 lanes
    if (valid) {
        const size_t i = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
        float dist = -1.0f;
        if (h.data != 0) {  // primary.comp.glsl:58-62
            is_hit = 1;
            gb.albedo[i] = h.data;
            gb.normal[i] = normal_rgba8(h.face);
            gb.position[i] = make_float4(ceilf(h.hx) / 8.0f, ceilf(h.hy) / 8.0f, ceilf(h.hz) / 8.0f, 1.0f);
            if (HITBUF) {
                const float ex = h.hx / 8.0f - cam.pos[0], ey = h.hy / 8.0f - cam.pos[1], ez = h.hz / 8.0f - cam.pos[2];
                dist = sqrtf(ex * ex + ey * ey + ez * ez);
            }
        } else {  // :63-68
            float r, g, b;
            sky_dome2((zmask & 1u) ? 0.0f : dx, (zmask & 2u) ? 0.0f : dy, (zmask & 4u) ? 0.0f : dz, r, g, b);
            gb.albedo[i] = pack_rgba8(r, g, b, 1.0f);
            gb.normal[i] = 0xFFFFFFFFu;
            gb.position[i] = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
        }
        if (HITBUF) store_hit(gb.hit, i, h, dist);
    }
    if (COUNT) {
        warp_add(&counters->rays, valid ? 1u : 0u);
        warp_add(&counters->t_in, tc.t_in);
        warp_add(&counters->t_chunk, tc.t_chunk);
        warp_add(&counters->t_block, tc.t_block);
        warp_add(&counters->hits, is_hit);
    }
}

// shadow ray of one pixel: secondary.comp.glsl:36-50.  Returns the illumination texel.
template <class World, int COUNT>
__device__ __forceinline__ uint32_t shadow_pixel(const World &w, bool active, const ViewDev &v, float posx, float posy, float posz,
                                                 uint32_t normal, TripCounts &tc, uint32_t &hit) {
    const float nx = unorm8_to_float(normal & 255u), ny = unorm8_to_float((normal >> 8) & 255u), nz = unorm8_to_float((normal >> 16) & 255u);
    const float ox = posx + nx * 0.001f, oy = posy + ny * 0.001f, oz = posz + nz * 0.001f;
    Hit h;
    // (the sun clearance map is built for the step cap in force: uvt.cu rebuilds it when the cap grows)
    trace<World, COUNT, true>(w, active, ox, oy, oz, UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, (int)v.max_steps, (int)(8u * v.map_dim), h, tc);  // all 32 lanes
    if (!active) { hit = 0; return 0u; }
    hit = h.data != 0;
    bool shadowed = h.data != 0;
    if (v.entities && !shadowed) {  // an entity hit only matters when the terrain ray missed (same -0.3 either way)
        const float ex = ox - h.hx / 8.0f, ey = oy - h.hy / 8.0f, ez = oz - h.hz / 8.0f;
        shadowed = trace_entities(ox, oy, oz, UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, sqrtf(ex * ex + ey * ey + ez * ez));
    }
    // vec4(SUN_DIR, -/+0.3) -> RGBA8: (192,168,192,0) or (192,168,192,77)
    return pack_rgba8(UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, shadowed ? -0.3f : 0.3f);
}

// ---- secondary pass ------------------------------------------------------------------------
template <class World, int COUNT>
__global__ void __launch_bounds__(kTileThreads, UVT_MIN_BLOCKS) secondary_kernel(WorldArgs<World> wa, ViewDev v, GBufDev gb, DevCounters *counters) {
    __shared__ uint32_t s_masks[(UVT_SMEM_MASKS && kIsCompact<World>) ? kSmemMaskMats * 16 : 1];
    World w = wa.w;
    if constexpr (kIsCompact<World>) {
        stage_masks(s_masks, wa.masks, wa.n_mats);
        w.smem_masks = s_masks;
    }
    uint32_t x, ly, y;
    tile_pixel(x, ly);
    ly += v.row0;
    const bool valid = v.global_row(ly, y) && x < v.W;
    TripCounts tc = {0, 0, 0};
    uint32_t traced = 0, early = 0, hit = 0;
    const size_t i = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
    float4 pos = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
    uint32_t nrm = 0;
    if (valid) {
        pos = gb.position[i];
        nrm = gb.normal[i];
    }
    const bool shoot = valid && !(pos.x < 0.0f || pos.y < 0.0f || pos.z < 0.0f);  // :26-29
    const uint32_t il = shadow_pixel<World, COUNT>(w, shoot, v, pos.x, pos.y, pos.z, nrm, tc, hit);  // all 32 lanes
    if (valid) {
        gb.illum[i] = shoot ? il : 0u;
        traced = shoot ? 1u : 0u;
        early = shoot ? 0u : 1u;
    }
    if (COUNT) {
        warp_add(&counters->rays, traced);
        warp_add(&counters->t_in, tc.t_in);
        warp_add(&counters->t_chunk, tc.t_chunk);
        warp_add(&counters->t_block, tc.t_block);
        warp_add(&counters->hits, hit);
        warp_add(&counters->early_out, early);
    }
}

// SkyDome2 once more, for the blit only: the shaded frame is held to the reference within 1/255 per channel, which
// leaves room for a reciprocal square root and plain multiplications (relative error ~1e-6) where the primary pass's sky colour
// keeps the exactly rounded divisions.
__device__ __forceinline__ void sky_dome2_fast(float rx, float ry, float rz, float &r, float &g, float &b) {
    const float inv_sl = 0.7991607f;  // 1 / |SUN_DIR|
    const float inv_rl = rsqrtf(rx * rx + ry * ry + rz * rz);
    const float dot = (UVT_SUN_X * rx + UVT_SUN_Y * ry + UVT_SUN_Z * rz) * (inv_sl * inv_rl);
    const float sun = fminf(fmaxf(dot, 0.0f), 1.2f);
    const float s2 = sun * sun, s4 = s2 * s2, p8 = s4 * s4, p3 = s2 * sun;
    const float k = ry * 0.2f;
    r = 0.675f - k + 0.4f * p8 + 0.2f * p3;
    g = 0.785f - k * 0.5f + 0.24f * p8 + 0.08f * p3;
    b = 0.825f - k + 0.04f * p8 + 0.04f * p3;
}

// blit.fragment.glsl:23-36 for one pixel.  Within the 1/255 tolerance of the frame: texel bytes become floats by a multiplication,
// the vignette's pow runs through exp2 / log2 (__powf), the lit term uses sky_dome2_fast and is skipped where it is multiplied by an
// illumination alpha of 0 (shadowed pixels: `color += 0 * SkyDome2(..)`).  The crosshair predicate stays exact — a pixel flipping
// in or out of it would change by half its value.
__device__ __forceinline__ uint32_t shade_pixel(const ViewDev &v, uint32_t x, uint32_t y, uint32_t albedo, uint32_t illum) {
    const float tx = ((float)x + 0.5f) / (float)v.W, ty = ((float)y + 0.5f) / (float)v.H;
    const float k255 = 1.0f / 255.0f;
    float c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = (float)((albedo >> (8 * k)) & 255u) * k255;
    if ((illum >> 24) != 0u) {
        const float ia = (float)(illum >> 24) * k255;
        float r, g, b;
        sky_dome2_fast((float)(illum & 255u) * k255, (float)((illum >> 8) & 255u) * k255, (float)((illum >> 16) & 255u) * k255, r, g, b);
        c[0] += ia * r; c[1] += ia * g; c[2] += ia * b; c[3] += ia;
    }
    const float cx = tx - 0.5f, cy = ty - 0.5f;
    if (fabsf(cx) <= 0.0021f && fabsf(cy) <= 0.0021f && sqrtf(cx * cx + cy * cy) <= 0.002f) {  // :33 mix(color, (1,1,1,0.4), 0.5)
        c[0] = c[0] * 0.5f + 0.5f; c[1] = c[1] * 0.5f + 0.5f;
        c[2] = c[2] * 0.5f + 0.5f; c[3] = c[3] * 0.5f + 0.2f;
    }
    const float vx = tx * (1.0f - tx), vy = ty * (1.0f - ty);
    const float grad = __powf(vx * vy * 15.0f, 0.6f * 0.3f);
    return pack_rgba8(grad * c[0], grad * c[1], grad * c[2], grad * c[3]);
}

// Frame target: either the ctx's own compact band buffer or caller memory (possibly a
// peer-mapped pointer on another GPU) addressed by GLOBAL row.
struct FrameTarget {
    uint32_t *ptr;
    uint32_t global_rows;  // 1: index by global row y (full-frame target); 0: by local row
};

__global__ void __launch_bounds__(256) shade_kernel(ViewDev v, GBufDev gb, FrameTarget ft) {
    const uint32_t x = blockIdx.x * 64u + (threadIdx.x & 63u);
    const uint32_t ly = v.row0 + blockIdx.y * 4u + (threadIdx.x >> 6);
    uint32_t y;
    if (!v.global_row(ly, y) || x >= v.W) return;
    const size_t i = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
    const uint32_t out = shade_pixel(v, x, y, gb.albedo[i], gb.illum[i]);
    if (ft.global_rows) ft.ptr[(size_t)y * v.W + x] = out;
    else ft.ptr[i] = out;
}

// ---- secondary pass + blit of the same pixel in one launch (uvt_dispatch_frame) ---------------------
// The illumination texel is stored as always (it is an output image of the reference), but the blit takes it from the register
// instead of reading it back, and one launch with its tail is gone.  Same shadow_pixel, same shade_pixel: identical pixels.
template <class World>
__global__ void __launch_bounds__(kTileThreads, UVT_MIN_BLOCKS) secondary_shade_kernel(WorldArgs<World> wa, ViewDev v, GBufDev gb, FrameTarget ft) {
    __shared__ uint32_t s_masks[(UVT_SMEM_MASKS && kIsCompact<World>) ? kSmemMaskMats * 16 : 1];
    World w = wa.w;
    if constexpr (kIsCompact<World>) {
        stage_masks(s_masks, wa.masks, wa.n_mats);
        w.smem_masks = s_masks;
    }
    uint32_t x, ly, y;
    tile_pixel(x, ly);
    ly += v.row0;
    const bool valid = v.global_row(ly, y) && x < v.W;
    TripCounts tc = {0, 0, 0};
    uint32_t hit = 0;
    const size_t i = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
    float4 pos = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
    uint32_t nrm = 0;
    if (valid) {
        pos = gb.position[i];
        nrm = gb.normal[i];
    }
    const bool shoot = valid && !(pos.x < 0.0f || pos.y < 0.0f || pos.z < 0.0f);  // secondary.comp.glsl:26-29
    const uint32_t il = shadow_pixel<World, 0>(w, shoot, v, pos.x, pos.y, pos.z, nrm, tc, hit);  // all 32 lanes
    if (!valid) return;
    const uint32_t illum = shoot ? il : 0u;
    gb.illum[i] = illum;
    const uint32_t out = shade_pixel(v, x, y, gb.albedo[i], illum);
    if (ft.global_rows) ft.ptr[(size_t)y * v.W + x] = out;
    else ft.ptr[i] = out;
}

// ---- fused frame: primary + secondary + shade in one launch -------------------------------
// Results are identical to the three separate passes: the shadow ray starts from the same
// quantised position/normal the G-buffer would hold.
template <class World, bool GBUF, bool BATCH>
__global__ void __launch_bounds__(kTileThreads, UVT_MIN_BLOCKS_FRAME) frame_kernel(WorldArgs<World> wa, const CamDev *__restrict__ cams, CamDev cam0, ViewDev v,
                                                         uint32_t shadow_steps, GBufDev gb, FrameTarget ft) {
    __shared__ uint32_t s_masks[(UVT_SMEM_MASKS && kIsCompact<World>) ? kSmemMaskMats * 16 : 1];
    World w = wa.w;
    if constexpr (kIsCompact<World>) {
        stage_masks(s_masks, wa.masks, wa.n_mats);
        w.smem_masks = s_masks;
    }
    uint32_t x, ly, y;
    tile_pixel(x, ly);
    const bool valid = v.global_row(ly, y) && x < v.W;
    const CamDev &cam = BATCH ? cams[blockIdx.z] : cam0;
    float dx = 0.0f, dy = 0.0f, dz = 1.0f, sx = 0.0f, sy = 0.0f, sz = 0.0f;
    if (valid) primary_ray(cam, v, x, y, dx, dy, dz, sx, sy, sz);
    const uint32_t zmask = (dx == 0.0f ? 1u : 0u) | (dy == 0.0f ? 2u : 0u) | (dz == 0.0f ? 4u : 0u);  // see primary_kernel
    if (zmask & 1u) dx = 0.001f;
    if (zmask & 2u) dy = 0.001f;
    if (zmask & 4u) dz = 0.001f;
    Hit h;
    TripCounts tc;
    trace<World, 0>(w, valid, sx, sy, sz, dx, dy, dz, (int)v.max_steps, (int)(8u * v.map_dim), h, tc);  // all 32 lanes
    const size_t i = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
    uint32_t albedo, normal, illum = 0u;
    float4 pos;
    if (h.data != 0) {
        albedo = h.data;
        normal = normal_rgba8(h.face);
        pos = make_float4(ceilf(h.hx) / 8.0f, ceilf(h.hy) / 8.0f, ceilf(h.hz) / 8.0f, 1.0f);
    } else {
        float r, g, b;
        sky_dome2((zmask & 1u) ? 0.0f : dx, (zmask & 2u) ? 0.0f : dy, (zmask & 4u) ? 0.0f : dz, r, g, b);
        albedo = pack_rgba8(r, g, b, 1.0f);
        normal = 0xFFFFFFFFu;
        pos = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
    }
    {
        const bool shoot = valid && !(pos.x < 0.0f || pos.y < 0.0f || pos.z < 0.0f);
        ViewDev vs = v;
        vs.max_steps = shadow_steps;
        uint32_t hit;
        const uint32_t il = shadow_pixel<World, 0>(w, shoot, vs, pos.x, pos.y, pos.z, normal, tc, hit);  // all 32 lanes
        if (shoot) illum = il;
    }
    if (!valid) return;
    if (GBUF) {
        gb.albedo[i] = albedo;
        gb.normal[i] = normal;
        gb.position[i] = pos;
        gb.illum[i] = illum;
    }
    const uint32_t out = shade_pixel(v, x, y, albedo, illum);
    if (ft.global_rows) ft.ptr[(size_t)y * v.W + x] = out;
    else ft.ptr[i] = out;
}

// terrain_edit.comp.glsl:10-17: the centre pick ray (rayUV = 0)
template <class World>
__global__ void pick_kernel(WorldArgs<World> wa, CamDev cam, ViewDev v, uint8_t *out_hit) {
    __shared__ uint32_t s_masks[(UVT_SMEM_MASKS && kIsCompact<World>) ? kSmemMaskMats * 16 : 1];
    World w = wa.w;
    if constexpr (kIsCompact<World>) {
        stage_masks(s_masks, wa.masks, wa.n_mats);
        w.smem_masks = s_masks;
    }
    const bool lead = threadIdx.x == 0;
    if (threadIdx.x >= 32) return;  // one warp traces; lane 0 carries the ray
    float q[4];
    for (int i = 0; i < 4; ++i) {
        float acc = cam.mat[0 * 4 + i] * 0.0f;
        acc = acc + cam.mat[1 * 4 + i] * 0.0f;
        acc = acc + cam.mat[2 * 4 + i] * 1.0f;
        acc = acc + cam.mat[3 * 4 + i] * 1.0f;
        q[i] = acc;
    }
    const float len = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float dx = q[0] / len, dy = q[1] / len, dz = q[2] / len;
    const float D = (float)v.map_dim;
    float tn, tf;
    intersect_aabb(cam.pos[0], cam.pos[1], cam.pos[2], dx, dy, dz, 0.0f, 0.0f, 0.0f, D, D, D, tn, tf);
    const float t0 = gmax(tn, 0.0f);
    Hit h;
    TripCounts tc;
    trace<World, 0>(w, lead, cam.pos[0] + dx * t0 - v.epsilon, cam.pos[1] + dy * t0 - v.epsilon, cam.pos[2] + dz * t0 - v.epsilon,
                    dx, dy, dz, (int)v.max_steps, (int)(8u * v.map_dim), h, tc);
    if (!lead) return;
    float dist = -1.0f;
    if (h.data != 0) {
        const float ex = h.hx / 8.0f - cam.pos[0], ey = h.hy / 8.0f - cam.pos[1], ez = h.hz / 8.0f - cam.pos[2];
        dist = sqrtf(ex * ex + ey * ey + ez * ez);
    }
    store_hit(out_hit, 0, h, dist);
}

// ---- the distinct block words of the committed bricks (the material table of the compact layout) --------------
// Open-addressed set in global memory; a world holds a few dozen distinct words, so almost every probe is a plain read of
// an entry that is already there.  `overflow` is set when the set fills up (more words than the 8-bit layout can name anyway).
constexpr uint32_t kWordSetSlots = 2048;
__global__ void distinct_words_kernel(const uint32_t *__restrict__ bricks, size_t n_words, uint32_t *table, uint32_t *overflow) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4u;
    uint32_t last = 0;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4u; i < n_words; i += stride) {
        const uint4 wv = *reinterpret_cast<const uint4 *>(bricks + i);
        const uint32_t wds[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t wd = wds[k];
            if (wd == 0u || wd == last) continue;
            last = wd;
            uint32_t slot = (wd * 2654435761u) >> 21;
            for (uint32_t probe = 0;; ++probe) {
                const uint32_t v = *(volatile uint32_t *)&table[slot];
                if (v == wd) break;
                if (v == 0u) {
                    const uint32_t old = atomicCAS(&table[slot], 0u, wd);
                    if (old == 0u || old == wd) break;
                }
                if (probe >= kWordSetSlots) { *overflow = 1u; break; }
                slot = (slot + 1u) & (kWordSetSlots - 1u);
            }
        }
    }
}

// ---- world repack: reference u32 bricks -> 8-bit material bricks ---------------------------
// mat_lut maps a block word to its material id through a small open-addressed table built on the host.
__global__ void repack_bricks_kernel(const uint32_t *__restrict__ bricks, uint8_t *__restrict__ bricks8, size_t n_words,
                                     const uint32_t *__restrict__ lut_keys, const uint8_t *__restrict__ lut_vals, uint32_t lut_mask) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 4u;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4u; i < n_words; i += stride) {
        const uint4 wv = *reinterpret_cast<const uint4 *>(bricks + i);
        const uint32_t wds[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t m = 0;
            if (wds[k] != 0) {
                uint32_t hsh = (wds[k] * 2654435761u) & lut_mask;
                while (lut_keys[hsh] != wds[k]) hsh = (hsh + 1u) & lut_mask;  // every word present was inserted by the host
                m = lut_vals[hsh];
            }
            packed |= m << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(bricks8 + i) = packed;
    }
}


// ---- chunk distance field for the fast path (trace_map_fast) ------------------------------
// dist[c] = Chebyshev distance (in chunks) from chunk c to the nearest NON-EMPTY chunk or to the
// outside of the map, capped at kFieldCap; separable min-max passes over x, y, z.
constexpr int kFieldCap = 32;
constexpr uint32_t kNoChunk = 0xFFFFFFFFu;  // brick_chunk[] of a brick slot no chunk uses

// The commit kernels below run over a box of chunks (the whole map for a full commit, the neighbourhood of an edit for
// an incremental one): thread t -> chunk (x, y, z) of the box, i = its linear index in the map.
struct ChunkBox {
    int ox, oy, oz, ex, ey, ez;
    __host__ __device__ size_t count() const { return (size_t)ex * ey * ez; }
};

__device__ __forceinline__ bool box_chunk(const ChunkBox &b, int cd, int &x, int &y, int &z, size_t &i) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.count()) return false;
    x = b.ox + (int)(t % b.ex);
    y = b.oy + (int)((t / b.ex) % b.ey);
    z = b.oz + (int)(t / ((size_t)b.ex * b.ey));
    i = (size_t)x + (size_t)cd * ((size_t)y + (size_t)z * cd);
    return true;
}

__global__ void field_pass_x_kernel(const uint32_t *__restrict__ chunks, uint8_t *__restrict__ out, int cd, ChunkBox box) {
    int x, y, z;
    size_t i;
    if (!box_chunk(box, cd, x, y, z, i)) return;
    int best = kFieldCap;
    if (chunks[i] != 0) best = 0;
    else {
        for (int k = 1; k < best; ++k) {
            const bool lo = (x - k < 0) || chunks[i - k] != 0;
            const bool hi = (x + k >= cd) || chunks[i + k] != 0;
            if (lo || hi) { best = k; break; }
        }
    }
    out[i] = (uint8_t)best;
}

// axis 1 = y (stride cd), axis 2 = z (stride cd^2); a value depends on inputs at most kFieldCap - 1 chunks away along the axis
__global__ void field_pass_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int cd, int axis, ChunkBox box) {
    int x, y, z;
    size_t i;
    if (!box_chunk(box, cd, x, y, z, i)) return;
    const int a = axis == 1 ? y : z;
    const size_t stride = axis == 1 ? (size_t)cd : (size_t)cd * cd;
    int best = in[i];
    for (int k = 1; k < best; ++k) {
        const int lo = (a - k < 0) ? 0 : (int)in[i - (size_t)k * stride];
        const int hi = (a + k >= cd) ? 0 : (int)in[i + (size_t)k * stride];
        const int v = max(k, min(lo, hi));
        best = min(best, v);
    }
    out[i] = (uint8_t)best;
}

// Empty chunks that touch a non-empty chunk (dist == 1) get a "virtual" brick: it holds no
// material, only per-block clearances, so that rays skimming the terrain also get free trips.
__global__ void count_virtual_kernel(const uint32_t *__restrict__ chunks, const uint8_t *__restrict__ dist, size_t n, unsigned int *counter) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool v = i < n && chunks[i] == 0 && dist[i] == 1;
    const unsigned int m = __ballot_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(counter, (unsigned int)__popc(m));
}

// chunks2[(cd+1)^3]: bit 31 = far-empty chunk with low byte n_free = 8*(dist-1) - 1; else 0-based brick
// index (real bricks keep the reference numbering, virtual bricks follow).  brick_chunk[b] = linear chunk index.
__global__ void build_chunks2_kernel(const uint32_t *__restrict__ chunks, const uint8_t *__restrict__ dist, uint32_t *__restrict__ chunks2,
                                     uint32_t *__restrict__ brick_chunk, int cd, uint32_t n_real, unsigned int *counter) {
    const int cd1 = cd + 1;
    const size_t n = (size_t)cd1 * cd1 * cd1;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % cd1), y = (int)((i / cd1) % cd1), z = (int)(i / ((size_t)cd1 * cd1));
    uint32_t e = 0x80000000u;  // guard layer: empty, nothing known about what follows
    if (x < cd && y < cd && z < cd) {
        const size_t j = (size_t)x + (size_t)cd * ((size_t)y + (size_t)z * cd);
        const uint32_t c = chunks[j];
        if (c != 0) {
            e = c - 1u;
            brick_chunk[e] = (uint32_t)j;
        } else if (dist[j] == 1) {
            e = n_real + atomicAdd(counter, 1u);
            brick_chunk[e] = (uint32_t)j;
        } else {
            const int r = (int)dist[j] - 1;  // rings of empty in-map chunks around this one (>= 1 here)
            e = 0x80000000u | (uint32_t)min(max(8 * r - 1, 0), 255);
        }
    }
    chunks2[i] = e;
}

// rowmask[b][z][y] = 8 x-occupancy bits of brick b (material bytes only; clearances are not written yet)
__global__ void brick_rowmask_kernel(const uint8_t *__restrict__ bricks8, size_t n_rows, uint8_t *__restrict__ rowmask) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const uint2 v = *reinterpret_cast<const uint2 *>(bricks8 + i * 8);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if ((v.x >> (8 * k)) & 0xFFu) m |= 1u << k;
        if ((v.y >> (8 * k)) & 0xFFu) m |= 1u << (k + 4);
    }
    rowmask[i] = (uint8_t)m;
}

// Column tops for sky_sealed(): tops32[x + dim*z] = highest occupied block y + 1 of the column (0: empty column).
// One CTA per real brick, one thread per (x, z) column of the brick; run on material bytes (before the clearances are written).
__global__ void __launch_bounds__(64) column_tops_kernel(const uint8_t *__restrict__ bricks8, const uint32_t *__restrict__ brick_chunk, int cd,
                                                         unsigned int *__restrict__ tops32) {
    const uint32_t b = blockIdx.x;
    const uint32_t cj = brick_chunk[b];
    if (cj == kNoChunk) return;  // a brick no chunk names
    const int cx = (int)(cj % cd), cy = (int)((cj / cd) % cd), cz = (int)(cj / ((uint32_t)cd * cd));
    const int lx = threadIdx.x & 7, lz = threadIdx.x >> 3;
    const int dim = cd * 8;
    for (int ly = 7; ly >= 0; --ly) {
        const uint8_t v = bricks8[(size_t)b * 512u + lx + 8 * ly + 64 * lz];
        if (v != 0 && v < kMatLimit) {
            atomicMax(&tops32[(size_t)(cx * 8 + lx) + (size_t)dim * (cz * 8 + lz)], (unsigned int)(cy * 8 + ly + 1));
            break;
        }
    }
}

// clear4[qx + (dim/4)*qz] = max of tops32 over the 4x4-block group grown by one block on every side
__global__ void quad_clear_kernel(const unsigned int *__restrict__ tops32, uint16_t *__restrict__ clear4, int dim) {
    const int qdim = dim >> 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= qdim * qdim) return;
    const int qx = i % qdim, qz = i / qdim;
    unsigned int m = 0;
    for (int z = max(4 * qz - 1, 0); z <= min(4 * qz + 4, dim - 1); ++z)
        for (int x = max(4 * qx - 1, 0); x <= min(4 * qx + 4, dim - 1); ++x) m = max(m, tops32[(size_t)x + (size_t)dim * z]);
    clear4[i] = (uint16_t)m;
}

// Next level of the column-tops pyramid: out[Q] = max of `in` over the 4x4 cells of group Q (the last groups may be partial).
// clear4 (4x4 blocks, grown by one block) -> clear16 -> clear64: line_free_trips walks the coarsest level first.
__global__ void coarse_clear_kernel(const uint16_t *__restrict__ in, uint16_t *__restrict__ out, int in_dim) {
    const int out_dim = (in_dim + 3) >> 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_dim * out_dim) return;
    const int cx = i % out_dim, cz = i / out_dim;
    unsigned int m = 0;
    for (int z = 4 * cz; z < min(4 * cz + 4, in_dim); ++z)
        for (int x = 4 * cx; x < min(4 * cx + 4, in_dim); ++x) m = max(m, (unsigned int)in[x + in_dim * z]);
    out[i] = (uint16_t)m;
}

// Sun clearance of the shadow pass, per block column.  Every shadow ray has the direction SUN_DIR = (s, u, s), all components
// positive, so how long it stays in open air depends only on where it starts.  Take a ray whose state lies in column (X, Z),
// block row r.  Travelling up every axis the state never moves down or back: `within += dir * t` adds non-negative numbers and
// a reset puts the stepped axis exactly on the face it reached, so in fp32, too, every later state has y >= the row's floor.
// A lookup names the state's block or, when a `within` component has rounded up to the step size, the next one up that axis —
// never a lower row, at most one column further in +x and / or +z: top2[c] = the highest column top (first all-empty row) over
// columns c + [0, 1]^2 covers every lookup made from column c.  The line runs along the diagonal of the x-z plane
// (s_x = s_z): it only visits columns (X + a, Z + b) with a, b >= 0 and |a - b| <= 1, and it is inside such a column only after
// an x- (or z-) advance of more than max(a, b) - 1 blocks, i.e. after climbing (u / s) * max(max(a, b) - 1, 0) blocks.  The
// real trajectory follows the ideal line to ~1e-5 block over 48 trips: 0.01 block of every credit is given up for that, and the
// columns (X + k + 1, Z + k - 1), (X + k - 1, Z + k + 1) a drifting state could clip at a cell corner are scanned too (with the
// credit of one step less).  Hence with
//     sun1[X,Z] = max over the scanned (a, b) of  ceil(top2[X + a, Z + b] - max(credit(a, b) - 0.01, 0))
// a state in a row >= sun1 looks up nothing but empty blocks for `steps` trips.  Columns beyond the x / z faces make the
// value infinite (the ray could leave the map before the cap); the top face is the caller's sun_row_max.  Reach: `steps`
// trips cover at most (steps + 5) / (2 s + u) of parameter, s times that along x.
// (x0, z0)-(x1, z1): the rectangle of columns to (re)compute, inclusive — everything, or the surroundings of an edit.
__global__ void top3_kernel(const unsigned int *__restrict__ tops32, uint16_t *__restrict__ top2, int dim, int x0, int z0, int x1, int z1) {
    const int w = x1 - x0 + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * (z1 - z0 + 1)) return;
    const int x = x0 + i % w, z = z0 + i / w;
    unsigned int m = 0;
    for (int zz = z; zz <= min(z + 1, dim - 1); ++zz)
        for (int xx = x; xx <= min(x + 1, dim - 1); ++xx) m = max(m, tops32[(size_t)xx + (size_t)dim * zz]);
    top2[(size_t)x + (size_t)dim * z] = (uint16_t)min(m, 0xFFFEu);
}

__host__ __device__ inline int sun_reach_columns(int steps) {
    return (int)(UVT_SUN_X * (float)(steps + 5) / (2.0f * UVT_SUN_X + UVT_SUN_Y)) + 2;
}

__global__ void sun_clear_kernel(const uint16_t *__restrict__ top2, uint16_t *__restrict__ sun1, int dim, int steps, int x0, int z0, int x1, int z1) {
    const int w = x1 - x0 + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * (z1 - z0 + 1)) return;
    const int x = x0 + i % w, z = z0 + i / w;
    const int K = sun_reach_columns(steps);
    const float climb = UVT_SUN_Y / UVT_SUN_X;
    int m = 0;
    bool open = false;
    for (int k = 0; k <= K; ++k) {
        const float credit = fmaxf(climb * (float)max(k - 1, 0) - 0.01f, 0.0f), credit_drift = fmaxf(climb * (float)max(k - 2, 0) - 0.01f, 0.0f);
        for (int j = 0; j < 5; ++j) {  // (a, b) = (k, k), (k, k - 1), (k - 1, k); corner drift: (k + 1, k - 1), (k - 1, k + 1)
            const int a = j == 2 ? k - 1 : (j == 3 ? k + 1 : (j == 4 ? k - 1 : k));
            const int b = j == 1 ? k - 1 : (j == 3 ? k - 1 : (j == 4 ? k + 1 : k));
            if (a < 0 || b < 0) continue;
            const int cx = x + a, cz = z + b;
            if (cx >= dim || cz >= dim) { open = true; continue; }
            m = max(m, (int)ceilf((float)top2[(size_t)cx + (size_t)dim * cz] - (j >= 3 ? credit_drift : credit)));
        }
    }
    sun1[(size_t)x + (size_t)dim * z] = open ? (uint16_t)0xFFFF : (uint16_t)min(m, 0xFFFE);
}

// Block-level clearance.  One CTA per brick, one thread per block.  The occupancy of the 5x5x5 chunk
// neighbourhood is staged in shared memory as 40x40 rows of 40 x-bits (out-of-map = occupied); each
// empty block searches growing Chebyshev shells for the nearest occupied block, D capped at 16, and
// stores kMatLimit + max(D - 2, 0) (see trace_map_fast for the bound).
constexpr int kClearCap = 16;

__device__ __forceinline__ void clearance_brick(const uint32_t *__restrict__ chunks2, int cd, uint32_t b, int cx, int cy, int cz,
                                                const uint8_t *__restrict__ rowmask, uint8_t *__restrict__ bricks8, unsigned long long *rows) {
    const int cd1 = cd + 1;
    for (int r = threadIdx.x; r < 40 * 40; r += blockDim.x) {
        const int yy = r % 40, zz = r / 40;                 // 0..39 -> block offset -16..23
        const int ncy = cy + (yy >> 3) - 2, ncz = cz + (zz >> 3) - 2;
        unsigned long long bits = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int ncx = cx + k - 2;
            uint32_t m8;
            if (ncx < 0 || ncy < 0 || ncz < 0 || ncx >= cd || ncy >= cd || ncz >= cd) m8 = 0xFFu;  // outside the map
            else {
                const uint32_t e = chunks2[(size_t)ncx + (size_t)cd1 * ((size_t)ncy + (size_t)ncz * cd1)];
                m8 = ((int)e < 0) ? 0u : rowmask[(size_t)e * 64u + (size_t)(zz & 7) * 8u + (size_t)(yy & 7)];
            }
            bits |= (unsigned long long)m8 << (8 * k);
        }
        rows[r] = bits;
    }
    __syncthreads();
    const int lx = threadIdx.x & 7, ly = (threadIdx.x >> 3) & 7, lz = threadIdx.x >> 6;
    const size_t addr = (size_t)b * 512u + threadIdx.x;
    const uint8_t cur = bricks8[addr];
    if (cur != 0 && cur < kMatLimit) return;  // material block (a stale clearance code is recomputed)
    const int x0 = lx + 16, y0 = ly + 16, z0 = lz + 16;
    int D = kClearCap;
    for (int r = 1; r < kClearCap; ++r) {
        const unsigned long long range = ((1ull << (2 * r + 1)) - 1ull) << (x0 - r);
        const unsigned long long ends = (1ull << (x0 - r)) | (1ull << (x0 + r));
        bool found = false;
        for (int dz = -r; dz <= r && !found; ++dz) {
            const bool zedge = dz == -r || dz == r;
            const unsigned long long *row = &rows[(z0 + dz) * 40 + y0];
            for (int dy = -r; dy <= r; ++dy) {
                const unsigned long long m = (zedge || dy == -r || dy == r) ? range : ends;  // only the new shell
                if (row[dy] & m) { found = true; break; }
            }
        }
        if (found) { D = r; break; }
    }
    bricks8[addr] = (uint8_t)(kMatLimit + max(D - 2, 0));
}

// all bricks (full commit): one CTA per brick slot
__global__ void __launch_bounds__(512) clearance_kernel(const uint32_t *__restrict__ chunks2, int cd, const uint32_t *__restrict__ brick_chunk,
                                                        const uint8_t *__restrict__ rowmask, uint8_t *__restrict__ bricks8) {
    __shared__ unsigned long long rows[40 * 40];  // [z'][y'], bit i = x' = i - 16 relative to the brick origin
    const uint32_t b = blockIdx.x;
    const uint32_t cj = brick_chunk[b];
    if (cj == kNoChunk) return;
    clearance_brick(chunks2, cd, b, (int)(cj % cd), (int)((cj / cd) % cd), (int)(cj / ((uint32_t)cd * cd)), rowmask, bricks8, rows);
}

// the bricks of a chunk box (incremental commit): grid = box extent, origin (ox, oy, oz)
__global__ void __launch_bounds__(512) clearance_box_kernel(const uint32_t *__restrict__ chunks2, int cd, int ox, int oy, int oz,
                                                            const uint8_t *__restrict__ rowmask, uint8_t *__restrict__ bricks8) {
    __shared__ unsigned long long rows[40 * 40];
    const int cx = ox + (int)blockIdx.x, cy = oy + (int)blockIdx.y, cz = oz + (int)blockIdx.z;
    const int cd1 = cd + 1;
    const uint32_t e = chunks2[(size_t)cx + (size_t)cd1 * ((size_t)cy + (size_t)cz * cd1)];
    if ((int)e < 0) return;  // far-empty chunk: no brick
    clearance_brick(chunks2, cd, e, cx, cy, cz, rowmask, bricks8, rows);
}

// Dense block grid for the traversal kernels: dense[x + dim*(z + dim*y)] = the brick byte of block (x, y, z);
// blocks of far-empty chunks get kMatLimit + min(n_free, 30) of their chunk.  One CTA per chunk, one thread per x-row.
__device__ __forceinline__ void dense_fill_chunk(const uint32_t *__restrict__ chunks2, const uint8_t *__restrict__ bricks8,
                                                 uint8_t *__restrict__ dense, int cd, int cx, int cy, int cz) {
    const int cd1 = cd + 1;
    const int ly = threadIdx.x & 7, lz = threadIdx.x >> 3;
    const uint32_t e = chunks2[(size_t)cx + (size_t)cd1 * ((size_t)cy + (size_t)cz * cd1)];
    uint2 v;
    if ((int)e < 0) {
        const uint32_t b = kMatLimit + min(e & 0xFFu, 30u);
        v.x = v.y = b * 0x01010101u;
    } else {
        v = *reinterpret_cast<const uint2 *>(bricks8 + (size_t)e * 512u + 8 * ly + 64 * lz);
    }
    const size_t dim = (size_t)cd * 8;
    *reinterpret_cast<uint2 *>(dense + (size_t)cx * 8 + dim * ((size_t)(cz * 8 + lz) + dim * (size_t)(cy * 8 + ly))) = v;
}

__global__ void __launch_bounds__(64) dense_fill_kernel(const uint32_t *__restrict__ chunks2, const uint8_t *__restrict__ bricks8,
                                                        uint8_t *__restrict__ dense, int cd) {
    const size_t cj = blockIdx.x;
    dense_fill_chunk(chunks2, bricks8, dense, cd, (int)(cj % cd), (int)((cj / cd) % cd), (int)(cj / ((size_t)cd * cd)));
}

__global__ void __launch_bounds__(64) dense_fill_box_kernel(const uint32_t *__restrict__ chunks2, const uint8_t *__restrict__ bricks8,
                                                            uint8_t *__restrict__ dense, int cd, int ox, int oy, int oz) {
    dense_fill_chunk(chunks2, bricks8, dense, cd, ox + (int)blockIdx.x, oy + (int)blockIdx.y, oz + (int)blockIdx.z);
}

__global__ void __launch_bounds__(64) dense_fill_list_kernel(const uint32_t *__restrict__ chunks2, const uint8_t *__restrict__ bricks8,
                                                             uint8_t *__restrict__ dense, int cd, const uint32_t *__restrict__ list) {
    const uint32_t cj = list[blockIdx.x];
    dense_fill_chunk(chunks2, bricks8, dense, cd, (int)(cj % cd), (int)((cj / cd) % cd), (int)(cj / ((uint32_t)cd * cd)));
}

// ---- incremental commit (uvt_world_commit_region) ---------------------------------
// Write the chunk entries of a box into the device table; *changed is set when any entry differs from the committed one.
__global__ void apply_chunk_box_kernel(uint32_t *__restrict__ chunks, const uint32_t *__restrict__ ents, int cd, int ox, int oy, int oz,
                                       int bx, int by, int bz, unsigned int *changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bx * by * bz) return;
    const int x = ox + i % bx, y = oy + (i / bx) % by, z = oz + i / (bx * by);
    const size_t j = (size_t)x + (size_t)cd * ((size_t)y + (size_t)z * cd);
    const uint32_t e = ents[i];
    if (chunks[j] != e) {
        chunks[j] = e;
        atomicOr(changed, 1u);
    }
}

// Repack listed bricks (u32 words -> material bytes) and refresh their row masks.  One CTA per brick, one thread per x-row.
__global__ void __launch_bounds__(64) repack_list_kernel(const uint32_t *__restrict__ bricks, uint8_t *__restrict__ bricks8, uint8_t *__restrict__ rowmask,
                                                         const uint32_t *__restrict__ list, const uint32_t *__restrict__ lut_keys,
                                                         const uint8_t *__restrict__ lut_vals, uint32_t lut_mask) {
    const size_t b = list[blockIdx.x];
    const size_t row = b * 64u + threadIdx.x;
    const uint4 lo = *reinterpret_cast<const uint4 *>(bricks + row * 8u), hi = *reinterpret_cast<const uint4 *>(bricks + row * 8u + 4u);
    const uint32_t wds[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t packed[2] = {0u, 0u}, mask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t m = 0;
        if (wds[k] != 0) {
            uint32_t hsh = (wds[k] * 2654435761u) & lut_mask;
            while (lut_keys[hsh] != wds[k]) hsh = (hsh + 1u) & lut_mask;
            m = lut_vals[hsh];
            mask |= 1u << k;
        }
        packed[k >> 2] |= m << (8 * (k & 3));
    }
    *reinterpret_cast<uint2 *>(bricks8 + row * 8u) = make_uint2(packed[0], packed[1]);
    rowmask[row] = (uint8_t)mask;
}

// empty chunks that newly touch a non-empty chunk and hold no virtual brick yet
__global__ void count_new_virtual_kernel(const uint32_t *__restrict__ chunks, const uint8_t *__restrict__ dist, const uint32_t *__restrict__ chunks2,
                                         int cd, uint32_t v_base, unsigned int *counter, ChunkBox box) {
    int x, y, z;
    size_t j;
    bool v = false;
    if (box_chunk(box, cd, x, y, z, j) && chunks[j] == 0 && dist[j] == 1) {
        const int cd1 = cd + 1;
        const uint32_t old = chunks2[(size_t)x + (size_t)cd1 * ((size_t)y + (size_t)z * cd1)];
        v = (int)old < 0 || old < v_base;
    }
    const unsigned int m = __ballot_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(counter, (unsigned int)__popc(m));
}

// chunks2 after a chunk-table change: real and existing virtual bricks keep their slots, new virtual bricks are
// appended (zeroed), far-empty distances are refreshed; every chunk whose entry changed is listed (dense refill).
__global__ void update_chunks2_kernel(const uint32_t *__restrict__ chunks, const uint8_t *__restrict__ dist, uint32_t *__restrict__ chunks2,
                                      uint32_t *__restrict__ brick_chunk, uint8_t *__restrict__ bricks8, uint8_t *__restrict__ rowmask, int cd,
                                      uint32_t v_base, unsigned int *n_virtual, uint32_t *__restrict__ changed_list, uint32_t changed_cap,
                                      unsigned int *changed_count, ChunkBox box) {
    int x, y, z;
    size_t j;
    if (!box_chunk(box, cd, x, y, z, j)) return;
    const int cd1 = cd + 1;
    const size_t i = (size_t)x + (size_t)cd1 * ((size_t)y + (size_t)z * cd1);
    const uint32_t old = chunks2[i];
    const uint32_t c = chunks[j];
    uint32_t e;
    if (c != 0) {
        e = c - 1u;
        brick_chunk[e] = (uint32_t)j;
    } else if (dist[j] == 1) {
        if ((int)old >= 0 && old >= v_base) e = old;
        else {
            e = v_base + atomicAdd(n_virtual, 1u);  // capacity was checked with count_new_virtual_kernel
            brick_chunk[e] = (uint32_t)j;
            uint4 *b8 = reinterpret_cast<uint4 *>(bricks8 + (size_t)e * 512u);
            for (int k = 0; k < 32; ++k) b8[k] = make_uint4(0u, 0u, 0u, 0u);
            uint4 *rm = reinterpret_cast<uint4 *>(rowmask + (size_t)e * 64u);
            for (int k = 0; k < 4; ++k) rm[k] = make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
        const int r = (int)dist[j] - 1;
        e = 0x80000000u | (uint32_t)min(max(8 * r - 1, 0), 255);
    }
    if (e != old) {
        chunks2[i] = e;
        const unsigned int k = atomicAdd(changed_count, 1u);
        if (k < changed_cap) changed_list[k] = (uint32_t)j;
    }
}

// column tops of the real bricks in the chunk columns [ox, ox+gridDim.x) x [oz, oz+gridDim.z), all chunk rows (gridDim.y = cd)
__global__ void __launch_bounds__(64) column_tops_box_kernel(const uint32_t *__restrict__ chunks, const uint8_t *__restrict__ bricks8, int cd,
                                                             int ox, int oz, unsigned int *__restrict__ tops32) {
    const int cx = ox + (int)blockIdx.x, cy = (int)blockIdx.y, cz = oz + (int)blockIdx.z;
    const uint32_t c = chunks[(size_t)cx + (size_t)cd * ((size_t)cy + (size_t)cz * cd)];
    if (c == 0) return;
    const size_t b = c - 1u;
    const int lx = threadIdx.x & 7, lz = threadIdx.x >> 3;
    const int dim = cd * 8;
    for (int ly = 7; ly >= 0; --ly) {
        const uint8_t v = bricks8[b * 512u + lx + 8 * ly + 64 * lz];
        if (v != 0 && v < kMatLimit) {
            atomicMax(&tops32[(size_t)(cx * 8 + lx) + (size_t)dim * (cz * 8 + lz)], (unsigned int)(cy * 8 + ly + 1));
            break;
        }
    }
}

// y_clear = max over clear4 (= max over the column tops: clear4 is a dilated max of them)
__global__ void max_clear_kernel(const uint16_t *__restrict__ clear4, int n, unsigned int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int v = i < n ? clear4[i] : 0u;
    v = __reduce_max_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31u) == 0 && v) atomicMax(out, v);
}

// ---- layout checksum (tests: an incremental commit must leave what a full commit builds) ----
// Position-keyed sums, independent of brick slot numbering: out[0] over the dense grid bytes (materials by block word), out[1] over
// the brick-path view (brick bytes by global block, far-empty entries by chunk), out[2] over clear4.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__global__ void __launch_bounds__(64) layout_checksum_kernel(const uint32_t *__restrict__ chunks2, const uint8_t *__restrict__ bricks8,
                                                             const uint8_t *__restrict__ dense, const uint32_t *__restrict__ mat_word, int cd,
                                                             unsigned long long *out) {
    // material bytes enter as their block word (ids depend on the order materials were first seen), clearances as they are
    auto val = [&](uint32_t v) -> unsigned long long { return v < kMatLimit ? (unsigned long long)mat_word[v] + 1ull : 0x100000000ull + v; };
    const size_t cj = blockIdx.x;
    const int cx = (int)(cj % cd), cy = (int)((cj / cd) % cd), cz = (int)(cj / ((size_t)cd * cd));
    const int cd1 = cd + 1;
    const int ly = threadIdx.x & 7, lz = threadIdx.x >> 3;
    const size_t dim = (size_t)cd * 8;
    const uint32_t e = chunks2[(size_t)cx + (size_t)cd1 * ((size_t)cy + (size_t)cz * cd1)];
    unsigned long long hd = 0, hb = 0;
    for (int lx = 0; lx < 8; ++lx) {
        const size_t g = (size_t)(cx * 8 + lx) + dim * ((size_t)(cz * 8 + lz) + dim * (size_t)(cy * 8 + ly));
        const unsigned long long key = mix64(g + 1) | 1ull;
        if (dense) hd += val(dense[g]) * key;
        if ((int)e >= 0) hb += val(bricks8[(size_t)e * 512u + lx + 8 * ly + 64 * lz]) * key;
    }
    if ((int)e < 0 && threadIdx.x == 0) hb += (unsigned long long)(e & 0xFFu) * (mix64(cj + 0x9e3779b97f4a7c15ull) | 1ull) + 0x5bd1e995ull;
    for (int o = 16; o > 0; o >>= 1) {
        hd += __shfl_xor_sync(0xFFFFFFFFu, hd, o);
        hb += __shfl_xor_sync(0xFFFFFFFFu, hb, o);
    }
    if ((threadIdx.x & 31u) == 0) {
        if (hd) atomicAdd(out + 0, hd);
        if (hb) atomicAdd(out + 1, hb);
    }
}

// position-keyed checksum of a u16 map (clear4, top3, sun1: `salt` keeps them apart), added into out[2]
__global__ void clear4_checksum_kernel(const uint16_t *__restrict__ clear4, int n, unsigned long long *out, unsigned long long salt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = i < n ? (unsigned long long)(clear4[i] + 1u) * (mix64((unsigned long long)i + salt) | 1ull) : 0ull;
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xFFFFFFFFu, h, o);
    if ((threadIdx.x & 31u) == 0 && h) atomicAdd(out + 2, h);
}

// ---- bandwidth probes (roofline denominators) ----------------------------------
__global__ void __launch_bounds__(256) l2_read_kernel(const uint4 *__restrict__ buf, size_t n_vec, int repeats, uint32_t *sink) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < repeats; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            const uint4 q = __ldcg(buf + i);  // cache-global: bypass L1 so every load is served by L2
            acc += q.x ^ q.y ^ q.z ^ q.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n_vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) dst[i] = src[i];
}

// assemble interleaved row bands gathered from n_parts ranks into one frame (rank 0, after the NCCL gather)
__global__ void __launch_bounds__(256) deinterleave_kernel(const uint32_t *__restrict__ gathered, uint32_t *__restrict__ frame,
                                                           uint32_t W, uint32_t H, uint32_t band_rows, uint32_t n_parts, uint32_t rows_per_part) {
    const uint32_t x = blockIdx.x * 256u + threadIdx.x;
    const uint32_t y = blockIdx.y;
    if (x >= W || y >= H) return;
    const uint32_t b = y / band_rows, part = b % n_parts, lb = b / n_parts;
    const uint32_t ly = lb * band_rows + (y - b * band_rows);
    frame[(size_t)y * W + x] = gathered[((size_t)part * rows_per_part + ly) * W + x];
}

}  // namespace uvt
