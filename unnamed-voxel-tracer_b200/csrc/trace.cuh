// trace.cuh — device-side traversal of the two-level voxel world (sm_100a).
//
// Semantics follow assets/shaders/map.glsl:83-168 (traceMap) trip for trip: the step
// SEQUENCE is part of the observable result (the fp32 residual `within` accumulates
// rounding along the path and the iteration cap counts loop trips), so the freedom taken
// here is only in WHAT IS FETCHED per trip and how rays are scheduled — never in which
// steps are taken.  Compile with --fmad=false: every fp32 op is individually rounded.
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace uvt {

// GLSL 4.50 §8.3 min/max (NaN behaviour differs from fminf/fmaxf)
__device__ __forceinline__ float gmin(float x, float y) { return y < x ? y : x; }
__device__ __forceinline__ float gmax(float x, float y) { return x < y ? y : x; }

struct Hit {
    uint32_t data;     // HitInfo.data (colour), 0 = miss
    float hx, hy, hz;  // HitInfo.hit_pos, sub-voxel units
    uint32_t px, py, pz;
    uint32_t block;
    uint32_t face;     // 1..6
    uint32_t trips;
    uint32_t exit_kind;  // 0 hit, 1 cap, 2 left the map
};

#ifndef UVT_SMEM_MASKS
#define UVT_SMEM_MASKS 0  // 1: experiment build that stages the occupancy masks in shared memory per CTA.  Measured 7 % SLOWER on every
                          // config (c3 1.647 vs 1.530 ms): the 2 KB in use stay L1-resident anyway, and the staging costs a barrier per CTA
#endif
constexpr int kSmemMaskMats = 64;  // (UVT_SMEM_MASKS build) masks of material ids < 64 are staged in shared memory, the rest read through L1

struct TripCounts {
    uint32_t t_in, t_chunk, t_block;
};

// ---- world views ---------------------------------------------------------------------

// The reference SSBO layout read verbatim (map.glsl:31-47,57-60): u32 chunk table,
// u32[512] bricks, RGBA8 models addressed by slot (= mdl & 32767, the 32x32x32 slot grid).
struct WorldRef {
    const uint32_t *__restrict__ chunks;
    const uint32_t *__restrict__ bricks;
    const uint32_t *__restrict__ models;  // [n_slots][512]
    uint32_t cd;                          // chunks per axis
    uint32_t n_slots;

    // returns the block word; `chunk_hit` reports chunk entry != 0
    __device__ __forceinline__ uint32_t block_at(uint32_t px, uint32_t py, uint32_t pz, bool &chunk_hit) const {
        const uint32_t bx = px >> 3, by = py >> 3, bz = pz >> 3;
        const uint32_t cx = bx >> 3, cy = by >> 3, cz = bz >> 3;
        chunk_hit = false;
        if (cx >= cd || cy >= cd || cz >= cd) return 0;  // unsigned compare covers < 0
        const uint32_t idx = __ldg(&chunks[cx + cd * (cy + cz * cd)]);
        if (idx == 0) return 0;
        chunk_hit = true;
        return __ldg(&bricks[(size_t)(idx - 1) * 512u + (bx & 7u) + (((bz & 7u) << 3) + (by & 7u)) * 8u]);
    }
    // returns the atlas texel (0 = empty sub-voxel)
    __device__ __forceinline__ uint32_t sub_at(uint32_t block, uint32_t px, uint32_t py, uint32_t pz) const {
        const uint32_t slot = block & 32767u;
        if (slot >= n_slots) return 0;
        return __ldg(&models[slot * 512u + (px & 7u) + ((py & 7u) << 3) + ((pz & 7u) << 6)]);
    }
    __device__ __forceinline__ bool sub_solid(uint32_t block, uint32_t px, uint32_t py, uint32_t pz, uint32_t &color) const {
        color = sub_at(block, px, py, pz);
        return color != 0;
    }
    __device__ __forceinline__ uint32_t block_word(uint32_t block) const { return block; }
};

constexpr uint32_t kMatLimit = 224;  // brick bytes >= kMatLimit encode empty blocks (kMatLimit + free trips, see trace_map_fast)

// B200 layout (DESIGN.md "Data layout in HBM"): u32 chunk table, 8-bit material bricks
// (512 B instead of 2 KiB), per-material 512-bit sub-voxel occupancy masks staged in shared
// memory, colours and block words resolved only at the hit.
struct WorldCompact {
    const uint32_t *__restrict__ chunks;    // reference encoding (0 = empty else brick+1); generic path only
    const uint32_t *__restrict__ chunks2;   // [(cd+1)^3] fast-path encoding, see trace_map_fast
    const uint8_t *__restrict__ bricks8;    // [slots][512], value = material id, or kMatLimit + free trips for an empty block
    const uint32_t *__restrict__ mat_word;  // [256] material id -> block word
    const uint32_t *__restrict__ mat_color; // [256][512] material id -> model texels
    const uint32_t *smem_masks;             // [kSmemMaskMats][16] shared-memory copy of the first occupancy masks
    const uint32_t *__restrict__ g_masks;   // [256][16] all occupancy masks (global)
    uint32_t cd;
    uint32_t cd1;                           // cd + 1: stride of chunks2 (one guard layer on the high side)
    uint32_t n_real_bricks;                 // bricks [0, n_real) mirror reference bricks; the rest only carry clearances
    int32_t y_clear;                        // every block with y >= y_clear is empty (max occupied block y + 1)
    int32_t dim;                            // MAP_DIMENSION in blocks
    const uint8_t *__restrict__ dense;      // [dim^3] one byte per block, x + dim*(z + dim*y): the brick bytes (far-empty chunks:
                                            // kMatLimit + min(n_free, 30)) without the chunk indirection; nullptr when not built
    const uint16_t *__restrict__ clear64;   // [ceil(dim/64)^2] maximum of clear4 over 64x64-block column groups (coarsest level of the walk)
    const uint16_t *__restrict__ clear16;   // [ceil(dim/16)^2] ... over 16x16-block column groups
    const uint16_t *__restrict__ sun1;      // [dim^2] sun clearance per block column: a SUN_DIR ray whose state lies in this column, in a block
                                            // row AT OR ABOVE this value, meets nothing but empty in-map blocks for sun_steps trips (sun_clear_kernel)
    int32_t sun_row_max;                    // ... provided its block row is at most this (the ray stays under the top face)
    const uint16_t *__restrict__ clear4;    // [(dim/4)^2] per 4x4-block column group, grown by one block on every side:
                                            // every block with y >= clear4 there is empty (sky_sealed)

    __device__ __forceinline__ uint32_t block_at(uint32_t px, uint32_t py, uint32_t pz, bool &chunk_hit) const {
        const uint32_t bx = px >> 3, by = py >> 3, bz = pz >> 3;
        const uint32_t cx = bx >> 3, cy = by >> 3, cz = bz >> 3;
        chunk_hit = false;
        if (cx >= cd || cy >= cd || cz >= cd) return 0;
        const uint32_t idx = __ldg(&chunks[cx + cd * (cy + cz * cd)]);
        if (idx == 0) return 0;
        chunk_hit = true;
        const uint32_t b8 = __ldg(&bricks8[(size_t)(idx - 1) * 512u + (bx & 7u) + (((bz & 7u) << 3) + (by & 7u)) * 8u]);
        return b8 < kMatLimit ? b8 : 0u;  // empty blocks carry their clearance
    }
    __device__ __forceinline__ bool sub_solid(uint32_t mat, uint32_t px, uint32_t py, uint32_t pz, uint32_t &color) const {
        const uint32_t bit = (px & 7u) + ((py & 7u) << 3) + ((pz & 7u) << 6);
        const uint32_t word = mask_word(mat, bit);
        if (((word >> (bit & 31u)) & 1u) == 0) return false;
        color = __ldg(&mat_color[mat * 512u + bit]);
        return true;
    }
    __device__ __forceinline__ uint32_t block_word(uint32_t mat) const { return __ldg(&mat_word[mat]); }
    __device__ __forceinline__ uint32_t mask_word(uint32_t mat, uint32_t bit) const {
#if UVT_SMEM_MASKS
        return mat < (uint32_t)kSmemMaskMats ? smem_masks[mat * 16u + (bit >> 5)] : __ldg(&g_masks[mat * 16u + (bit >> 5)]);
#else
        return __ldg(&g_masks[mat * 16u + (bit >> 5)]);
#endif
    }
};

// Same data; selects (at compile time) the traversal that reads the dense block grid instead of
// chunk table + bricks.
struct WorldDense : WorldCompact {};

template <class World>
constexpr bool kIsCompact = std::is_same<World, WorldCompact>::value || std::is_same<World, WorldDense>::value;

// ---- traceMap ------------------------------------------------------------------------
// map.glsl:83-168.  `bound` = 8 * MAP_DIMENSION.
template <class World, bool COUNT>
__device__ __forceinline__ void trace_map(const World &w, float ox, float oy, float oz, float dx, float dy, float dz,
                                          int max_steps, int bound, Hit &out, TripCounts &tc) {
    if (dx == 0.0f) dx = 0.001f;  // :85-90
    if (dy == 0.0f) dy = 0.001f;
    if (dz == 0.0f) dz = 0.001f;

    // raySign / rayPositivity / rayInv: :94-96 (no zero components remain)
    const bool posx = dx > 0.0f, posy = dy > 0.0f, posz = dz > 0.0f;
    const float invx = 1.0f / dx, invy = 1.0f / dy, invz = 1.0f / dz;

    int mi = 0;  // :98
    const float o8x = ox * 8.0f, o8y = oy * 8.0f, o8z = oz * 8.0f;
    int gx = __float2int_rz(o8x), gy = __float2int_rz(o8y), gz = __float2int_rz(o8z);  // :101
    float wx = o8x - (float)gx, wy = o8y - (float)gy, wz = o8z - (float)gz;             // :102
    bool big = false;  // stepSize == 3

    out.data = 0;
    out.hx = out.hy = out.hz = -1.0f;  // :167
    out.px = out.py = out.pz = 0xFFFFFFFFu;
    out.block = 0;
    out.face = 0;
    out.exit_kind = 1;
    if (COUNT) tc.t_in = tc.t_chunk = tc.t_block = 0;

    int trip = 0;
    for (; trip < max_steps; ++trip) {
        // :107 — one unsigned compare per axis covers both < 0 and >= bound
        if ((unsigned)gx >= (unsigned)bound || (unsigned)gy >= (unsigned)bound || (unsigned)gz >= (unsigned)bound) {
            // (the trip count is stored here and the function left at once: reading `trip` after a `break` out of this
            // loop was observed to yield a wrong count with nvcc 12.9 -O3 on one ray of a 60-pose sweep)
            out.exit_kind = 2;
            out.trips = (uint32_t)trip;
            return;
        }
        if (COUNT) tc.t_in++;
        const uint32_t px = (uint32_t)gx + __float2uint_rz(wx);  // :108
        const uint32_t py = (uint32_t)gy + __float2uint_rz(wy);
        const uint32_t pz = (uint32_t)gz + __float2uint_rz(wz);

        bool chunk_hit;
        const uint32_t block = w.block_at(px, py, pz, chunk_hit);  // :114
        if (COUNT && chunk_hit) tc.t_chunk++;

        if (block != 0) {
            if (COUNT) tc.t_block++;
            uint32_t color;
            if (w.sub_solid(block, px, py, pz, color)) {  // :117-118
                out.data = color;
                out.face = mi == 0 ? (posx ? 1u : 2u) : (mi == 1 ? (posy ? 3u : 4u) : (posz ? 5u : 6u));  // :119-125
                out.hx = (float)gx + wx;  // :127
                out.hy = (float)gy + wy;
                out.hz = (float)gz + wz;
                out.px = px; out.py = py; out.pz = pz;
                out.block = w.block_word(block);
                out.exit_kind = 0;
                out.trips = (uint32_t)trip + 1u;
                return;
            }
            if (big) {  // :131-135
                gx += __float2int_rz(wx);
                gy += __float2int_rz(wy);
                gz += __float2int_rz(wz);
                wx = wx - floorf(wx);
                wy = wy - floorf(wy);
                wz = wz - floorf(wz);
                big = false;
            }
        } else if (!big) {  // :140-144
            wx += (float)(gx & 7);
            wy += (float)(gy & 7);
            wz += (float)(gz & 7);
            gx &= ~7;
            gy &= ~7;
            gz &= ~7;
            big = true;
        }

        // dda stepping: :157-162
        const float stepf = big ? 8.0f : 1.0f;
        const float tx = ((posx ? stepf : 0.0f) - wx) * invx;
        const float ty = ((posy ? stepf : 0.0f) - wy) * invy;
        const float tz = ((posz ? stepf : 0.0f) - wz) * invz;
        mi = tx < ty ? (tx < tz ? 0 : 2) : (ty < tz ? 1 : 2);
        const float tm = mi == 0 ? tx : (mi == 1 ? ty : tz);
        const int istep = big ? 8 : 1;
        wx += dx * tm;
        wy += dy * tm;
        wz += dz * tm;
        const float reset = stepf * 0.999f;  // float((1 - pos) << step) * 0.999f
        if (mi == 0) { gx += posx ? istep : -istep; wx = posx ? 0.0f : reset; }
        else if (mi == 1) { gy += posy ? istep : -istep; wy = posy ? 0.0f : reset; }
        else { gz += posz ? istep : -istep; wz = posz ? 0.0f : reset; }
    }
    out.trips = (uint32_t)max(max_steps, 0);  // iteration cap reached (:167)
}

// ---- conservative "nothing ahead" test for climbing rays -------------------------------------
// A ray with dir.y > 0 whose straight line stays at least kSkyMargin (one block) above every occupied block of the
// 4x4-block column groups it passes over (each grown by one block sideways, clear4) cannot meet a
// non-empty block: the reference DDA follows that line to within a fraction of a block (0.999 resets,
// fp32 rounding) and looks up `pos` at most one block beside it.  The test walks the column groups with
// a 2-D DDA until the line has covered more blocks (L1) than the remaining trips can cross, or has risen
// above the whole world.  It decides only WHETHER lookups can be skipped, never a hit, so it needs no
// bit-exact arithmetic.  (px, py, pz) in blocks, d = the ray direction with zero components patched.
#ifndef UVT_SKY_MARGIN
#define UVT_SKY_MARGIN 1.0f
#endif
// Blocks of air demanded between the line and the tops under it.  What is needed: the state of a climbing ray is
// the line's point at its parameter, exact in y (resets of an axis travelled upwards are exact) and up to 0.001 block
// per trip (the 0.999 resets) ahead along the horizontal axes; a lookup names the block of that point or, when `within`
// undershoots zero, the one behind it — within one block sideways (the groups are grown by one block) and never a row
// below the line's.  So y_in >= top would do in exact arithmetic; one block covers the fp32 slack of this test with room.
constexpr float kSkyMargin = UVT_SKY_MARGIN;

// CELL = blocks per column group: 4 (clear4) or 64 (clear64, the maximum of clear4 over 16x16 groups).
template <int CELL>
__device__ __forceinline__ bool sky_walk(const uint16_t *__restrict__ tops, int qdim, float y_all, float px, float py, float pz,
                                         float dx, float dy, float dz, float inv_dx, float inv_dz, float t_stop) {
    int qx = (int)(px * (1.0f / CELL)), qz = (int)(pz * (1.0f / CELL));
    const int sx = dx > 0.0f ? 1 : -1, sz = dz > 0.0f ? 1 : -1;
    float tmx = ((float)((qx + (dx > 0.0f ? 1 : 0)) * CELL) - px) * inv_dx;  // line parameter at the next x / z group boundary
    float tmz = ((float)((qz + (dz > 0.0f ? 1 : 0)) * CELL) - pz) * inv_dz;
    const float tdx = (float)CELL * fabsf(inv_dx), tdz = (float)CELL * fabsf(inv_dz);
    float t = 0.0f;
    for (int it = 0; it < 96; ++it) {
        if ((unsigned)qx >= (unsigned)qdim || (unsigned)qz >= (unsigned)qdim) return false;
        const float y_in = py + dy * t;  // lowest height of the line inside this group (it climbs)
        if (y_in - kSkyMargin < (float)__ldg(&tops[qx + qdim * qz])) return false;
        if (y_in >= y_all) return true;  // above every occupied block of the world
        if (tmx < tmz) { t = tmx; tmx += tdx; qx += sx; }
        else { t = tmz; tmz += tdz; qz += sz; }
        if (t > t_stop) return true;
    }
    return false;
}

// The coarse groups are tried first (a handful of cells for the whole reach: rays well above the terrain), the 4x4
// groups only when that fails.
__device__ __noinline__ bool sky_sealed(const uint16_t *__restrict__ clear4, const uint16_t *__restrict__ clear64, int dim, int y_clear,
                                        float px, float py, float pz, float dx, float dy, float dz, int trips_left) {
    const float inv_dx = 1.0f / dx, inv_dz = 1.0f / dz;
    // each trip crosses one block boundary, so after n trips the line has covered about n blocks in L1
    const float t_stop = (float)(trips_left + 4) / (fabsf(dx) + fabsf(dy) + fabsf(dz));
    const float y_all = (float)y_clear + kSkyMargin;
    if (sky_walk<64>(clear64, (dim + 63) >> 6, y_all, px, py, pz, dx, dy, dz, inv_dx, inv_dz, t_stop)) return true;
    return sky_walk<4>(clear4, dim >> 2, y_all, px, py, pz, dx, dy, dz, inv_dx, inv_dz, t_stop);
}

// ---- free trips from the column tops: "how long does this ray stay in open air?" ---------------------
// Generalises sky_sealed() to rays of ANY direction and to a trip COUNT instead of a yes/no: returns n, the number of
// FOLLOWING trips (after the current one, whose lookup found an empty block) that provably stay inside the map and look
// up empty blocks; n == n_rem means the ray is sealed (iteration-cap miss, map.glsl:167).
//
// Why the column tops decide this.  While every lookup is empty the ray stays at block steps, and its state S = g + within
// moves along straight pieces parallel to dir: `within += dir * t[minIdx]` advances all three axes by the same parameter,
// and the stepped axis is put exactly on the block face it reached (travelling up an axis) or 0.001 block past it
// (travelling down: the 0.999 reset, map.glsl:162).  So after any number of trips S = P + dir * tau + J with P the
// state now, tau >= 0 and J_k in [-0.001 * trips, 0] on the axes travelled downwards, 0 on the others.  A lookup names
// block floor(S) or, per axis, the next one up when `within` has rounded up to the step size (the carry of
// trace_map_fast's free-trip bound).  Hence, for the line point L = P + dir * tau:
//   * the looked-up COLUMN is within one block of L's column: clear4 / clear64 hold the tops of column groups grown by
//     one block sideways, so the group that contains L covers it;
//   * the looked-up ROW is >= floor(L.y - 0.001 * trips): it is empty when L.y - margin >= top of that group, with
//     margin = 0.001 * trips + slack for the fp32 arithmetic of THIS test (positions < 4096 blocks: errors < 0.003);
//   * g stays inside the map while L stays one block away from the three map faces ahead (t_face below).
// A trip crosses one block face, so after j trips |S - P|_1 < j + 3 blocks and tau < (j + 3.6) / |dir|_1: trips
// j < t_safe * |dir|_1 - 3.6 look up blocks the walk has covered, where [0, t_safe) is the parameter range proven
// clear.  The walk is a 2-D DDA over the 64x64-block groups first (a handful of cells for the whole reach) and continues
// over the 4x4-block groups from where the coarse level stopped.  It decides only WHETHER fetches can be skipped — the
// ray's own arithmetic is untouched — so it needs no bit-exact arithmetic (fused multiply-adds are fine here).
template <int CELL>
__device__ __forceinline__ float tops_walk(const uint16_t *__restrict__ tops, int qdim, float y_all, float margin, float t0, float t_end,
                                           float px, float py, float pz, float dx, float dy, float dz, float invx, float invz, int max_iter) {
    const float x0 = __fmaf_rn(dx, t0, px), z0 = __fmaf_rn(dz, t0, pz);
    int qx = (int)(x0 * (1.0f / CELL)), qz = (int)(z0 * (1.0f / CELL));
    const bool up = dy > 0.0f;
    const int sx = dx > 0.0f ? 1 : -1, sz = dz > 0.0f ? 1 : -1;
    float tmx = ((float)((qx + (dx > 0.0f ? 1 : 0)) * CELL) - px) * invx;  // line parameter at the next x / z group boundary
    float tmz = ((float)((qz + (dz > 0.0f ? 1 : 0)) * CELL) - pz) * invz;
    const float tdx = (float)CELL * fabsf(invx), tdz = (float)CELL * fabsf(invz);
    float t = t0;
    for (int it = 0; it < max_iter; ++it) {
        if ((unsigned)qx >= (unsigned)qdim || (unsigned)qz >= (unsigned)qdim) break;  // (t < t_face keeps the line inside: insurance)
        const float t_out = fminf(fminf(tmx, tmz), t_end);
        const float y_lo = __fmaf_rn(dy, up ? t : t_out, py) - margin;  // lowest point of the line inside this group
        if (y_lo < (float)__ldg(&tops[qx + qdim * qz])) break;
        if (up && y_lo >= y_all) return t_end;  // climbing above every occupied block of the world
        t = t_out;
        if (t >= t_end) break;
        if (tmx < tmz) { tmx += tdx; qx += sx; }
        else { tmz += tdz; qz += sz; }
    }
    return t;
}

__device__ __noinline__ int line_free_trips(const uint16_t *__restrict__ clear4, const uint16_t *__restrict__ clear16,
                                            const uint16_t *__restrict__ clear64, int dim, int y_clear,
                                            float px, float py, float pz, float dx, float dy, float dz, float invx, float invy, float invz,
                                            int n_rem) {
    const float l1 = fabsf(dx) + fabsf(dy) + fabsf(dz);
    const float t_want = (float)(n_rem + 5) / l1;
    // the line stays one block away from the map faces ahead for parameters below t_face
    const float hi = (float)(dim - 1);
    const float ax = (dx > 0.0f ? hi - px : px - 1.0f) * fabsf(invx);
    const float ay = (dy > 0.0f ? hi - py : py - 1.0f) * fabsf(invy);
    const float az = (dz > 0.0f ? hi - pz : pz - 1.0f) * fabsf(invz);
    const float t_end = fminf(t_want, fminf(fminf(ax, ay), az));
    if (!(t_end > 0.0f)) return 0;
    const float margin = 0.0625f + 0.001f * (float)(n_rem + 1);
    const float y_all = (float)y_clear;
    float t_safe = tops_walk<64>(clear64, (dim + 63) >> 6, y_all, margin, 0.0f, t_end, px, py, pz, dx, dy, dz, invx, invz, 12);
    if (t_safe < t_end) t_safe = tops_walk<16>(clear16, (dim + 15) >> 4, y_all, margin, t_safe, t_end, px, py, pz, dx, dy, dz, invx, invz, 32);
    if (t_safe < t_end) t_safe = tops_walk<4>(clear4, dim >> 2, y_all, margin, t_safe, t_end, px, py, pz, dx, dy, dz, invx, invz, 64);
    const int n = (int)(t_safe * l1 - 4.5f);
    return min(max(n, 0), n_rem);
}

// One DDA step (map.glsl:157-162), hand-scheduled and branch-free: 6 ops for t, 3 compares, min3,
// 6 ops for within += dir * t[minIdx], then the stepped axis is reset / advanced under its predicate.
// .rn ops are never contracted.  t[minIdx] is the minimum of the three (ties carry equal values;
// the fast path has no NaNs), and minIdx follows the reference's strict-less-than cascade.
// operands: %0-%5 = gx gy gz wx wy wz (read-write); then isx isy isz tgx tgy tgz invx invy invz dx dy dz rsx rsy rsz
#ifndef UVT_DDA_F32X2
#define UVT_DDA_F32X2 1  // x and y of `t` and of `within` through Blackwell's packed fp32 ops (FADD2 / FMUL2): 20 instead of 23 instructions per trip
#endif
#if UVT_DDA_F32X2
// add / sub / mul.rn.f32x2 round each half on its own exactly like the scalar ops.  The products dir * t[minIdx] stay scalar on
// purpose: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (one rounding) even under -fmad=false, which would break parity.
__device__ __forceinline__ void dda_step(int &gx, int &gy, int &gz, float &wx, float &wy, float &wz, int isx, int isy, int isz,
                                         float tgx, float tgy, float tgz, float invx, float invy, float invz,
                                         float dx, float dy, float dz, float rsx, float rsy, float rsz) {
    asm volatile("{\n\t"
                 ".reg .pred pxy, pmx, pmy, pmz;\n\t"
                 ".reg .f32 tx, ty, tz, tm, ax, ay, az;\n\t"
                 ".reg .b64 wxy, tgxy, invxy, txy, axy;\n\t"
                 "mov.b64 wxy, {%3, %4};\n\t"
                 "mov.b64 tgxy, {%9, %10};\n\t"
                 "mov.b64 invxy, {%12, %13};\n\t"
                 "sub.rn.f32x2 txy, tgxy, wxy;\n\t"
                 "sub.rn.f32 tz, %11, %5;\n\t"
                 "mul.rn.f32x2 txy, txy, invxy;\n\t"
                 "mul.rn.f32 tz, tz, %14;\n\t"
                 "mov.b64 {tx, ty}, txy;\n\t"
                 "setp.lt.f32 pxy, tx, ty;\n\t"
                 "setp.lt.and.f32 pmx, tx, tz, pxy;\n\t"
                 "setp.lt.and.f32 pmy, ty, tz, !pxy;\n\t"
                 "min.f32 tm, tx, ty;\n\t"
                 "min.f32 tm, tm, tz;\n\t"
                 "mul.rn.f32 ax, %15, tm;\n\t"
                 "mul.rn.f32 ay, %16, tm;\n\t"
                 "mul.rn.f32 az, %17, tm;\n\t"
                 "mov.b64 axy, {ax, ay};\n\t"
                 "add.rn.f32x2 wxy, wxy, axy;\n\t"
                 "add.rn.f32 %5, %5, az;\n\t"
                 "mov.b64 {%3, %4}, wxy;\n\t"
                 "or.pred pmz, pmx, pmy;\n\t"
                 "@pmx mov.f32 %3, %18;\n\t"
                 "@pmy mov.f32 %4, %19;\n\t"
                 "@!pmz mov.f32 %5, %20;\n\t"
                 "@pmx add.s32 %0, %0, %6;\n\t"
                 "@pmy add.s32 %1, %1, %7;\n\t"
                 "@!pmz add.s32 %2, %2, %8;\n\t"
                 "}"
                 : "+r"(gx), "+r"(gy), "+r"(gz), "+f"(wx), "+f"(wy), "+f"(wz)
                 : "r"(isx), "r"(isy), "r"(isz), "f"(tgx), "f"(tgy), "f"(tgz), "f"(invx), "f"(invy), "f"(invz),
                   "f"(dx), "f"(dy), "f"(dz), "f"(rsx), "f"(rsy), "f"(rsz));
}

// same step; additionally %6 %7 = (minIdx == 0), (minIdx == 1)
__device__ __forceinline__ void dda_step_last(int &gx, int &gy, int &gz, float &wx, float &wy, float &wz, int &mxi, int &myi,
                                              int isx, int isy, int isz, float tgx, float tgy, float tgz, float invx, float invy, float invz,
                                              float dx, float dy, float dz, float rsx, float rsy, float rsz) {
    asm volatile("{\n\t"
                 ".reg .pred pxy, pmx, pmy, pmz;\n\t"
                 ".reg .f32 tx, ty, tz, tm, ax, ay, az;\n\t"
                 ".reg .b64 wxy, tgxy, invxy, txy, axy;\n\t"
                 "mov.b64 wxy, {%3, %4};\n\t"
                 "mov.b64 tgxy, {%11, %12};\n\t"
                 "mov.b64 invxy, {%14, %15};\n\t"
                 "sub.rn.f32x2 txy, tgxy, wxy;\n\t"
                 "sub.rn.f32 tz, %13, %5;\n\t"
                 "mul.rn.f32x2 txy, txy, invxy;\n\t"
                 "mul.rn.f32 tz, tz, %16;\n\t"
                 "mov.b64 {tx, ty}, txy;\n\t"
                 "setp.lt.f32 pxy, tx, ty;\n\t"
                 "setp.lt.and.f32 pmx, tx, tz, pxy;\n\t"
                 "setp.lt.and.f32 pmy, ty, tz, !pxy;\n\t"
                 "min.f32 tm, tx, ty;\n\t"
                 "min.f32 tm, tm, tz;\n\t"
                 "mul.rn.f32 ax, %17, tm;\n\t"
                 "mul.rn.f32 ay, %18, tm;\n\t"
                 "mul.rn.f32 az, %19, tm;\n\t"
                 "mov.b64 axy, {ax, ay};\n\t"
                 "add.rn.f32x2 wxy, wxy, axy;\n\t"
                 "add.rn.f32 %5, %5, az;\n\t"
                 "mov.b64 {%3, %4}, wxy;\n\t"
                 "or.pred pmz, pmx, pmy;\n\t"
                 "@pmx mov.f32 %3, %20;\n\t"
                 "@pmy mov.f32 %4, %21;\n\t"
                 "@!pmz mov.f32 %5, %22;\n\t"
                 "@pmx add.s32 %0, %0, %8;\n\t"
                 "@pmy add.s32 %1, %1, %9;\n\t"
                 "@!pmz add.s32 %2, %2, %10;\n\t"
                 "selp.s32 %6, 1, 0, pmx;\n\t"
                 "selp.s32 %7, 1, 0, pmy;\n\t"
                 "}"
                 : "+r"(gx), "+r"(gy), "+r"(gz), "+f"(wx), "+f"(wy), "+f"(wz), "=r"(mxi), "=r"(myi)
                 : "r"(isx), "r"(isy), "r"(isz), "f"(tgx), "f"(tgy), "f"(tgz), "f"(invx), "f"(invy), "f"(invz),
                   "f"(dx), "f"(dy), "f"(dz), "f"(rsx), "f"(rsy), "f"(rsz));
}
#else
__device__ __forceinline__ void dda_step(int &gx, int &gy, int &gz, float &wx, float &wy, float &wz, int isx, int isy, int isz,
                                         float tgx, float tgy, float tgz, float invx, float invy, float invz,
                                         float dx, float dy, float dz, float rsx, float rsy, float rsz) {
    asm volatile("{\n\t"
                 ".reg .pred pxy, pmx, pmy, pmz;\n\t"
                 ".reg .f32 tx, ty, tz, tm, ax, ay, az;\n\t"
                 "sub.rn.f32 tx, %9, %3;\n\t"
                 "sub.rn.f32 ty, %10, %4;\n\t"
                 "sub.rn.f32 tz, %11, %5;\n\t"
                 "mul.rn.f32 tx, tx, %12;\n\t"
                 "mul.rn.f32 ty, ty, %13;\n\t"
                 "mul.rn.f32 tz, tz, %14;\n\t"
                 "setp.lt.f32 pxy, tx, ty;\n\t"
                 "setp.lt.and.f32 pmx, tx, tz, pxy;\n\t"
                 "setp.lt.and.f32 pmy, ty, tz, !pxy;\n\t"
                 "min.f32 tm, tx, ty;\n\t"
                 "min.f32 tm, tm, tz;\n\t"
                 "mul.rn.f32 ax, %15, tm;\n\t"
                 "mul.rn.f32 ay, %16, tm;\n\t"
                 "mul.rn.f32 az, %17, tm;\n\t"
                 "add.rn.f32 %3, %3, ax;\n\t"
                 "add.rn.f32 %4, %4, ay;\n\t"
                 "add.rn.f32 %5, %5, az;\n\t"
                 "or.pred pmz, pmx, pmy;\n\t"
                 "@pmx mov.f32 %3, %18;\n\t"
                 "@pmy mov.f32 %4, %19;\n\t"
                 "@!pmz mov.f32 %5, %20;\n\t"
                 "@pmx add.s32 %0, %0, %6;\n\t"
                 "@pmy add.s32 %1, %1, %7;\n\t"
                 "@!pmz add.s32 %2, %2, %8;\n\t"
                 "}"
                 : "+r"(gx), "+r"(gy), "+r"(gz), "+f"(wx), "+f"(wy), "+f"(wz)
                 : "r"(isx), "r"(isy), "r"(isz), "f"(tgx), "f"(tgy), "f"(tgz), "f"(invx), "f"(invy), "f"(invz),
                   "f"(dx), "f"(dy), "f"(dz), "f"(rsx), "f"(rsy), "f"(rsz));
}

// same step; additionally %6 %7 = (minIdx == 0), (minIdx == 1)
__device__ __forceinline__ void dda_step_last(int &gx, int &gy, int &gz, float &wx, float &wy, float &wz, int &mxi, int &myi,
                                              int isx, int isy, int isz, float tgx, float tgy, float tgz, float invx, float invy, float invz,
                                              float dx, float dy, float dz, float rsx, float rsy, float rsz) {
    asm volatile("{\n\t"
                 ".reg .pred pxy, pmx, pmy, pmz;\n\t"
                 ".reg .f32 tx, ty, tz, tm, ax, ay, az;\n\t"
                 "sub.rn.f32 tx, %11, %3;\n\t"
                 "sub.rn.f32 ty, %12, %4;\n\t"
                 "sub.rn.f32 tz, %13, %5;\n\t"
                 "mul.rn.f32 tx, tx, %14;\n\t"
                 "mul.rn.f32 ty, ty, %15;\n\t"
                 "mul.rn.f32 tz, tz, %16;\n\t"
                 "setp.lt.f32 pxy, tx, ty;\n\t"
                 "setp.lt.and.f32 pmx, tx, tz, pxy;\n\t"
                 "setp.lt.and.f32 pmy, ty, tz, !pxy;\n\t"
                 "min.f32 tm, tx, ty;\n\t"
                 "min.f32 tm, tm, tz;\n\t"
                 "mul.rn.f32 ax, %17, tm;\n\t"
                 "mul.rn.f32 ay, %18, tm;\n\t"
                 "mul.rn.f32 az, %19, tm;\n\t"
                 "add.rn.f32 %3, %3, ax;\n\t"
                 "add.rn.f32 %4, %4, ay;\n\t"
                 "add.rn.f32 %5, %5, az;\n\t"
                 "or.pred pmz, pmx, pmy;\n\t"
                 "@pmx mov.f32 %3, %20;\n\t"
                 "@pmy mov.f32 %4, %21;\n\t"
                 "@!pmz mov.f32 %5, %22;\n\t"
                 "@pmx add.s32 %0, %0, %8;\n\t"
                 "@pmy add.s32 %1, %1, %9;\n\t"
                 "@!pmz add.s32 %2, %2, %10;\n\t"
                 "selp.s32 %6, 1, 0, pmx;\n\t"
                 "selp.s32 %7, 1, 0, pmy;\n\t"
                 "}"
                 : "+r"(gx), "+r"(gy), "+r"(gz), "+f"(wx), "+f"(wy), "+f"(wz), "=r"(mxi), "=r"(myi)
                 : "r"(isx), "r"(isy), "r"(isz), "f"(tgx), "f"(tgy), "f"(tgz), "f"(invx), "f"(invy), "f"(invz),
                   "f"(dx), "f"(dy), "f"(dz), "f"(rsx), "f"(rsy), "f"(rsz));
}

#endif  // UVT_DDA_F32X2

// ---- traceMap, B200 fast path ----------------------------------------------------------
// Same trips, same arithmetic as trace_map (map.glsl:83-168); what changes is what is FETCHED
// and how the loop is laid out:
//
//  * FREE TRIPS.  chunks2[(cd+1)^3] marks far-empty chunks with bit 31 and carries n_free in its
//    low byte; every other chunk (non-empty, or empty but touching a non-empty one) owns a brick
//    of one byte per block: < kMatLimit = material id, >= kMatLimit = empty with
//    n_free = byte - kMatLimit.  n_free is the number of FOLLOWING trips that provably (a) stay
//    inside the map and (b) look up an empty block, so they run the DDA arithmetic only — no
//    bounds test, no `pos`, no loads.  Proof: at 8-sub-voxel steps g moves exactly one block along
//    one axis per trip and `pos` is in g's block or, when a `within` component rounds up to 8, one
//    block further; with D the Chebyshev distance (blocks) from the looked-up block to the nearest
//    non-empty block or map face: with c_j in {0,1}^3 the round-up carry of trip j, the block looked
//    up at trip j is P_0 - c_0 + (j unit steps) + c_j, within j + 1 of P_0, so trips j <= D - 2 are
//    free.  The chunk-level field gives D >= 8R + 1 for R empty chunk rings, hence n_free = 8R - 1
//    there; the block-level field (clearance_kernel) gives n_free = min(D, 16) - 2 inside bricks.
//    (COUNT == 1 ignores n_free so that the reference counters stay exact.)
//  * WARP LOCKSTEP.  All 32 lanes are at the same trip index.  After a round of lookups every live
//    lane knows how many of its next trips are free; the warp runs the minimum of those as a
//    divergence-free, branch-free DDA loop with a uniform trip count, then the lanes whose free
//    trips ran out look up again.  Lanes whose ray has ended keep executing the arithmetic on dead
//    state (no memory traffic) and are parked with limit = kDead.
//  * SEALED RAYS.  A ray whose remaining trips are all provably empty AND inside the map ends as the
//    reference's iteration-cap miss (data 0, trips = maxSteps) whatever its arithmetic would have
//    been, so it is retired at once: (a) trip + 1 + n_free >= maxSteps, or (b) the ray climbs
//    (dir.y > 0), its block row is above every occupied block of the world, and no map face in its
//    direction of travel is within maxSteps + 2 blocks (one block per trip at most), or (c) the ray
//    climbs and sky_sealed() proves its line stays a block above the terrain it passes over.
//    (c) is tried by the whole warp together at trips 4, 32, 64, 128 (a failed test is cheap).  Sky
//    and sun-shadow rays stop marching as soon as they clear the terrain.  Only a lookup that found an EMPTY block can seal
//    (the current trip's own block must still be tested).  (COUNT == 1 never seals: exact counters.)
//  * the guard layer of chunks2 (index cd on any axis) removes the chunk-range test: `pos` can
//    exceed the map by at most one block on the high side while g is in bounds.
//  * sub-voxel occupancy is a bit test in shared memory; colour and block word are fetched once,
//    at the hit.  Per-phase constants (target face, reset value, signed step per axis) live in
//    registers and are rewritten only when the step size changes.
//
// Rays with non-finite reciprocals or origins beyond 2^20 sub-voxels take the generic path,
// whose corner-case behaviour (NaN ordering, saturation) is the specification.
constexpr int kDead = 0x40000000;    // `limit` of a lane without a live ray
#ifndef UVT_SHORT_LOOKUP
#define UVT_SHORT_LOOKUP 1
#endif
#ifndef UVT_WALK_GAP
#define UVT_WALK_GAP 16
#endif
#ifndef UVT_WALK_BACKOFF_SHIFT
#define UVT_WALK_BACKOFF_SHIFT 2
#endif
#ifndef UVT_WALK_MIN_CLEAR
#define UVT_WALK_MIN_CLEAR 1
#endif
constexpr int kWalkMinClear = UVT_WALK_MIN_CLEAR;  // walk only when the block clearance of the lookup grants at least this many free trips
constexpr int kWalkUseful = 4;                  // a walk that proves fewer free trips than this counts as failed
constexpr int kWalkGap = UVT_WALK_GAP;          // trips between the end of a proven run and the lane's next walk
constexpr int kWalkBackoffShift = UVT_WALK_BACKOFF_SHIFT;  // ... after a failed walk: max_steps >> this (48 of 192 trips)

// Must be called by ALL 32 lanes of a warp (it uses full-mask warp reductions); `active` = this lane has a ray.
// SUN: the direction is camera.glsl's SUN_DIR (the shadow pass): the precomputed sun clearance (sun1) seals the ray at the
// first lookup above it, and the column-tops walk is not needed.
template <int COUNT, bool DENSE, bool SUN>
__device__ __forceinline__ void trace_map_fast(const WorldCompact &w, bool active, float ox, float oy, float oz, float dx, float dy, float dz,
                                               int max_steps, int bound, Hit &out, TripCounts &tc) {
    if (dx == 0.0f) dx = 0.001f;
    if (dy == 0.0f) dy = 0.001f;
    if (dz == 0.0f) dz = 0.001f;
    const float invx = 1.0f / dx, invy = 1.0f / dy, invz = 1.0f / dz;
    const float o8x = ox * 8.0f, o8y = oy * 8.0f, o8z = oz * 8.0f;
    const bool sane = fabsf(invx) < 1e30f && fabsf(invy) < 1e30f && fabsf(invz) < 1e30f &&
                      fabsf(dx) < 1e30f && fabsf(dy) < 1e30f && fabsf(dz) < 1e30f &&
                      fabsf(o8x) < 1048576.0f && fabsf(o8y) < 1048576.0f && fabsf(o8z) < 1048576.0f;
    const bool generic = active && (!sane || max_steps <= 0);
    if (generic) {  // rare lanes: the generic loop is the specification for corner cases
        trace_map<WorldCompact, COUNT == 1>(w, ox, oy, oz, dx, dy, dz, max_steps, bound, out, tc);
        if (COUNT == 2) tc.t_in = tc.t_chunk = tc.t_block = 0;
    }
    const bool fast = active && !generic;
    const bool posx = dx > 0.0f, posy = dy > 0.0f, posz = dz > 0.0f;

    int gx = __float2int_rz(o8x), gy = __float2int_rz(o8y), gz = __float2int_rz(o8z);
    float wx = o8x - (float)gx, wy = o8y - (float)gy, wz = o8z - (float)gz;

    // phase constants, stepSize 0
    float tgx = posx ? 1.0f : 0.0f, tgy = posy ? 1.0f : 0.0f, tgz = posz ? 1.0f : 0.0f;          // float(rayPositivity << step)
    float rsx = posx ? 0.0f : 0.999f, rsy = posy ? 0.0f : 0.999f, rsz = posz ? 0.0f : 0.999f;    // float((1 - pos) << step) * 0.999f
    int isx = posx ? 1 : -1, isy = posy ? 1 : -1, isz = posz ? 1 : -1;                           // raySign << step
    bool big = false;

    if (!generic) {
        out.data = 0;
        out.hx = out.hy = out.hz = -1.0f;
        out.px = out.py = out.pz = 0xFFFFFFFFu;
        out.block = 0;
        out.face = 0;
        out.exit_kind = 1;
        out.trips = 0;
        if (COUNT) tc.t_in = tc.t_chunk = tc.t_block = 0;
    }

    const uint32_t cd1 = w.cd1;
    int trip = 0;                     // warp-uniform
    int limit = fast ? 0 : kDead;     // trips in [trip, limit) need no lookup; kDead parks the lane
    bool mx = true, my = false;       // minIdx of the previous trip == 0 / == 1 (starts at 0, map.glsl:98)
    // the lane may run the column-tops walk (line_free_trips) from this trip on: at once, then again some trips after the
    // run it proved has been used up (kWalkGap), later after a walk that proved next to nothing (kWalkBackoff)
    int walk_at = COUNT == 1 ? kDead : 0;

    uint32_t cmat = 0;  // material of the block looked up last (0: none) — sub-voxel steps mostly stay inside it

    for (;;) {
        // ---- lookups (map.glsl:107-144) --------------------------------------------------
        // A round is paid for by the whole warp, so EVERY live lane looks up, not only the ones whose
        // free trips ran out: a lane still inside its free run re-reads an (empty) block and refreshes
        // its clearance from the new position, which keeps the lanes' lookups aligned (fewer rounds).
        bool slow = limit < kDead;
        bool walk = false;  // this round's lookup found an empty block and the lane is due for a column-tops walk
        if (UVT_SHORT_LOOKUP && DENSE && COUNT != 1 && slow) {
            // the common lookup, kept short: a block step (no round-up carry) inside the map that finds an empty block and
            // does not seal the ray.  Anything else falls through to the general code below, which redoes the lookup.
            const uint32_t gmax = max(max((uint32_t)gx, (uint32_t)gy), (uint32_t)gz);
            const bool inside = gmax < (uint32_t)bound;
            const uint32_t udim = (uint32_t)w.dim;
            uint32_t code = 0;
            if (inside) code = __ldg(&w.dense[(size_t)((uint32_t)gx >> 3) + (size_t)udim * (((uint32_t)gz >> 3) + udim * ((uint32_t)gy >> 3))]);
            const int lim0 = max(limit, trip + 1 + (int)code - (int)kMatLimit);
            const bool simple = big && code >= kMatLimit && fmaxf(fmaxf(wx, wy), wz) < 8.0f;
            if (simple && lim0 < max_steps) {  // (a run that reaches the cap seals the ray: general code)
                limit = lim0;
                slow = false;
                if (SUN) {  // within < 8: the state's block is g >> 3
                    const int row = gy >> 3;
                    if (row >= (int)__ldg(&w.sun1[((uint32_t)gx >> 3) + udim * ((uint32_t)gz >> 3)]) && row <= w.sun_row_max) {
                        out.trips = (uint32_t)max_steps;  // sealed: iteration-cap miss (map.glsl:167)
                        out.px = out.py = out.pz = 0xFFFFFFFFu;
                        limit = kDead;
                    }
                } else {
                    walk = trip >= walk_at && (int)code - (int)kMatLimit >= kWalkMinClear;
                }
#ifndef UVT_ROUND_STATS
                if (COUNT == 2) tc.t_in++;
#endif
            }
        }
        if (slow) {
            if ((unsigned)gx >= (unsigned)bound || (unsigned)gy >= (unsigned)bound || (unsigned)gz >= (unsigned)bound) {
                out.exit_kind = 2;
                out.trips = (uint32_t)trip;
                out.px = out.py = out.pz = 0xFFFFFFFFu;
                limit = kDead;
            } else {
#ifndef UVT_ROUND_STATS
                if (COUNT == 2) tc.t_in++;  // lookups performed
#endif
                uint32_t px = (uint32_t)gx, py = (uint32_t)gy, pz = (uint32_t)gz;
                uint32_t mat;
                int n_free = 0;
                if (DENSE && COUNT != 1) {
                    // dense block grid: at block steps with within < 8 the block of `pos` is g >> 3, so the
                    // float->int conversions of map.glsl:108 are only needed for sub-voxel steps / round-up carries
                    const bool exact = !big || fmaxf(fmaxf(wx, wy), wz) >= 8.0f;
                    uint32_t b8 = kMatLimit;  // `pos` one block past the high map face: empty, nothing known
                    const uint32_t udim = (uint32_t)w.dim;
                    if (exact) {
                        px += __float2uint_rz(wx);
                        py += __float2uint_rz(wy);
                        pz += __float2uint_rz(wz);
                    }
                    if (!exact || ((px >> 3) < udim && (py >> 3) < udim && (pz >> 3) < udim))
                        b8 = __ldg(&w.dense[(size_t)(px >> 3) + (size_t)udim * ((pz >> 3) + udim * (py >> 3))]);
                    const bool is_mat = b8 < kMatLimit;
                    n_free = is_mat ? 0 : (int)(b8 - kMatLimit);
                    mat = is_mat ? b8 : 0u;
                    if (is_mat && !exact) {  // the sub-voxel test needs the exact `pos`
                        px += __float2uint_rz(wx);
                        py += __float2uint_rz(wy);
                        pz += __float2uint_rz(wz);
                    }
                } else {
                    px += __float2uint_rz(wx);
                    py += __float2uint_rz(wy);
                    pz += __float2uint_rz(wz);
                    // out.p* hold the `pos` of the previous lookup: still inside that (non-empty) block?
                    if (!big && cmat != 0u && (((px ^ out.px) | (py ^ out.py) | (pz ^ out.pz)) < 8u)) {
                        mat = cmat;
                        if (COUNT == 1) tc.t_chunk++;
                    } else {
                        const uint32_t e = __ldg(&w.chunks2[(px >> 6) + cd1 * ((py >> 6) + (pz >> 6) * cd1)]);
                        if ((int)e < 0) {
                            n_free = (int)(e & 0xFFu);
                            mat = 0u;
                        } else {
                            if (COUNT == 1 && e < w.n_real_bricks) tc.t_chunk++;
                            const uint32_t b8 = __ldg(&w.bricks8[e * 512u + (((px >> 3) & 7u) | (py & 0x38u) | ((pz & 0x38u) << 3))]);
                            const bool is_mat = b8 < kMatLimit;
                            n_free = is_mat ? 0 : (int)(b8 - kMatLimit);
                            mat = is_mat ? b8 : 0u;
                        }
                        cmat = mat;
                    }
                }
                out.px = px; out.py = py; out.pz = pz;
                if (COUNT == 1) n_free = 0;  // exact reference counters need every lookup
                limit = max(limit, trip + 1 + n_free);  // an earlier guarantee stays valid
                bool seal = COUNT != 1 && mat == 0u && limit >= max_steps;
                limit = min(limit, max_steps);
                if (SUN && COUNT != 1) {  // the state's sub-voxel is `pos` (within >= 0), so its block is pos >> 3
                    const uint32_t qd = (uint32_t)w.dim;
                    if (mat == 0u && (px >> 3) < qd && (pz >> 3) < qd) {
                        const int row = (int)(py >> 3);
                        seal = seal || (row >= (int)__ldg(&w.sun1[(px >> 3) + qd * (pz >> 3)]) && row <= w.sun_row_max);
                    }
                } else {
                    walk = mat == 0u && !seal && trip >= walk_at && n_free >= kWalkMinClear;
                }
                if (seal) {
                    // sealed: nothing but empty in-map blocks until the iteration cap (map.glsl:167)
                    out.trips = (uint32_t)max_steps;
                    out.px = out.py = out.pz = 0xFFFFFFFFu;
                    limit = kDead;
                } else if (mat != 0u) {  // a block: test the sub-voxel
                    if (COUNT == 1) tc.t_block++;
                    const uint32_t bit = (px & 7u) | ((py & 7u) << 3) | ((pz & 7u) << 6);
                    const uint32_t word = w.mask_word(mat, bit);
                    if ((word >> (bit & 31u)) & 1u) {
                        out.data = __ldg(&w.mat_color[mat * 512u + bit]);
                        out.face = mx ? (posx ? 1u : 2u) : (my ? (posy ? 3u : 4u) : (posz ? 5u : 6u));
                        out.hx = (float)gx + wx;
                        out.hy = (float)gy + wy;
                        out.hz = (float)gz + wz;
                        out.block = __ldg(&w.mat_word[mat]);
                        out.exit_kind = 0;
                        out.trips = (uint32_t)trip + 1u;
                        limit = kDead;
                    } else if (big) {  // drop to sub-voxel steps (map.glsl:131-135)
                        gx += __float2int_rz(wx);
                        gy += __float2int_rz(wy);
                        gz += __float2int_rz(wz);
                        wx = wx - floorf(wx);
                        wy = wy - floorf(wy);
                        wz = wz - floorf(wz);
                        big = false;
                        tgx = posx ? 1.0f : 0.0f; tgy = posy ? 1.0f : 0.0f; tgz = posz ? 1.0f : 0.0f;
                        rsx = posx ? 0.0f : 0.999f; rsy = posy ? 0.0f : 0.999f; rsz = posz ? 0.0f : 0.999f;
                        isx = posx ? 1 : -1; isy = posy ? 1 : -1; isz = posz ? 1 : -1;
                    }
                } else if (!big) {  // rise to block steps (map.glsl:140-144)
                    wx += (float)(gx & 7);
                    wy += (float)(gy & 7);
                    wz += (float)(gz & 7);
                    gx &= ~7;
                    gy &= ~7;
                    gz &= ~7;
                    big = true;
                    tgx = posx ? 8.0f : 0.0f; tgy = posy ? 8.0f : 0.0f; tgz = posz ? 8.0f : 0.0f;
                    rsx = posx ? 0.0f : 8.0f * 0.999f; rsy = posy ? 0.0f : 8.0f * 0.999f; rsz = posz ? 0.0f : 8.0f * 0.999f;
                    isx = posx ? 8 : -8; isy = posy ? 8 : -8; isz = posz ? 8 : -8;
                }
            }
        }

        // ---- column-tops walk: many free trips at once, or the seal (sealed rays, above) ----------
        if (COUNT != 1 && !SUN && walk) {
            const int n_rem = max_steps - trip - 1;
            const int n = line_free_trips(w.clear4, w.clear16, w.clear64, w.dim, w.y_clear, ((float)gx + wx) * 0.125f, ((float)gy + wy) * 0.125f,
                                          ((float)gz + wz) * 0.125f, dx, dy, dz, invx, invy, invz, n_rem);
            if (n >= n_rem) {  // nothing but empty in-map blocks until the iteration cap (map.glsl:167)
                out.trips = (uint32_t)max_steps;
                out.px = out.py = out.pz = 0xFFFFFFFFu;
                limit = kDead;
            } else {
                limit = max(limit, trip + 1 + n);
                walk_at = trip + 1 + n + (n < kWalkUseful ? max_steps >> kWalkBackoffShift : kWalkGap);
            }
        }

        // ---- how many trips can the whole warp run without a lookup? ----------------------
        // live lanes: 1 <= limit - trip <= max_steps - trip; parked lanes: huge
        const int k = __reduce_min_sync(0xFFFFFFFFu, limit - trip);
        if (k >= kDead / 2) break;  // no live lane left
#ifdef UVT_ROUND_STATS
        if (COUNT == 2) {  // experiment build: t_in = rounds, t_chunk = single-trip rounds forced by a sub-voxel lane, t_block = by a block-step lane
            const bool lim1 = limit - trip == 1;
            const unsigned sub = __ballot_sync(0xFFFFFFFFu, lim1 && !big), blk = __ballot_sync(0xFFFFFFFFu, lim1 && big);
            if ((threadIdx.x & 31u) == 0) {
                tc.t_in++;
                if (sub) tc.t_chunk++;
                else if (blk) tc.t_block++;
            }
        }
#endif

        // ---- k DDA steps, branch-free (map.glsl:157-162) -----------------------------------
#ifdef UVT_DDA_UNROLL  // experiment knob (tools/variants.py); the default leaves the unrolling to the compiler (x4)
#define UVT_PRAGMA_(x) _Pragma(#x)
#define UVT_PRAGMA(x) UVT_PRAGMA_(x)
        UVT_PRAGMA(unroll UVT_DDA_UNROLL)
#endif
        for (int j = 1; j < k; ++j) dda_step(gx, gy, gz, wx, wy, wz, isx, isy, isz, tgx, tgy, tgz, invx, invy, invz, dx, dy, dz, rsx, rsy, rsz);
        {   // the last step of the run also reports minIdx: a hit in the next lookup needs it for the face id
            int mxi, myi;
            dda_step_last(gx, gy, gz, wx, wy, wz, mxi, myi, isx, isy, isz, tgx, tgy, tgz, invx, invy, invz, dx, dy, dz, rsx, rsy, rsz);
            mx = mxi != 0;
            my = myi != 0;
        }
        trip += k;
        if (trip >= max_steps) {  // iteration cap: miss (map.glsl:167); in lockstep it ends every live lane at once
            if (limit < kDead) {
                out.trips = (uint32_t)trip;
                out.px = out.py = out.pz = 0xFFFFFFFFu;
            }
            break;
        }
    }
    if (COUNT == 1 && fast) tc.t_in = out.trips;  // every executed trip passed the bounds test
}

// dispatch: the compact world takes the fast path
// COUNT: 0 none, 1 exact reference counters (t_in, t_chunk, t_block), 2 fast-path statistics (t_in = lookups performed)
// Must be called by all 32 lanes of a warp; lanes without a ray pass active = false.
template <class World, int COUNT, bool SUN = false>
__device__ __forceinline__ void trace(const World &w, bool active, float ox, float oy, float oz, float dx, float dy, float dz,
                                      int max_steps, int bound, Hit &out, TripCounts &tc) {
    if constexpr (kIsCompact<World>) trace_map_fast<COUNT, std::is_same<World, WorldDense>::value, SUN>(w, active, ox, oy, oz, dx, dy, dz, max_steps, bound, out, tc);
    else if (active) {
        trace_map<World, COUNT == 1>(w, ox, oy, oz, dx, dy, dz, max_steps, bound, out, tc);
        if (COUNT == 2) tc.t_in = tc.t_chunk = tc.t_block = 0;
    }
}

}  // namespace uvt
