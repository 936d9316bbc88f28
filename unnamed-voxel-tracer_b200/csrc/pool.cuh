// pool.cuh — the B200 traversal scheduler: a per-CTA ray pool with phase-wise compaction.
//
// Why: with free trips and sealed rays (trace.cuh) a large share of rays ends long before the
// iteration cap — sky rays are sealed after a handful of trips, hit rays end wherever they hit — but
// in a pixel-per-thread kernel a warp stays busy until its LAST lane is done.  Here a CTA owns a
// 16x16 pixel tile (two rays per thread).  Rays are advanced in PHASES of trips
// [0,4) [4,16) [16,32) [32,64) [64,128) ...; between phases the rays still alive are compacted in
// shared memory (warp ballot + popc ranks, CTA prefix over the four warps) so that the next phase
// runs on densely packed warps again.  Finished rays write their pixel straight away.
//
// Inside a phase a warp runs the lockstep rounds of trace.cuh: all rays of the pool are at the same
// trip index at a phase boundary, so the trip counter stays uniform.  The sealed-ray test
// (sky_sealed) is tried at the phase boundaries by all candidate rays at once.
//
// Semantics are those of trace_map_fast (same trips, same arithmetic, same lookups); only the
// assignment of rays to lanes changes.  State that survives a phase boundary: direction, reciprocal,
// grid coordinates, within, limit, flags (step size, last minIdx, climbs, zero components, pixel).
#pragma once

#include "kernels.cuh"

namespace uvt {

constexpr int kPoolTile = 16;                 // CTA tile is kPoolTile x kPoolTile pixels
constexpr int kPoolRays = kPoolTile * kPoolTile;
constexpr int kPoolMaskMats = kSmemMaskMats;  // sub-voxel masks of material ids < 64 live in shared memory

// flags word of a pooled ray
constexpr uint32_t kFlagBig = 1u, kFlagMx = 2u, kFlagMy = 4u, kFlagClimbs = 8u, kFlagZeroShift = 4;  // bits 4-6: zero components of the unpatched direction

struct PoolRay {
    float dx, dy, dz, invx, invy, invz, wx, wy, wz;
    int gx, gy, gz;
    int limit;
    uint32_t flags;   // kFlag* | pixel id << 16
    uint32_t c0, c1;  // COUNT: t_chunk / t_block (COUNT == 1) or lookups (COUNT == 2)
};

template <int COUNT>
struct PoolWords { static constexpr int n = COUNT ? 16 : 14; };

template <int COUNT>
__device__ __forceinline__ void pool_store(uint32_t (*pool)[kPoolRays], int i, const PoolRay &r) {
    pool[0][i] = __float_as_uint(r.dx); pool[1][i] = __float_as_uint(r.dy); pool[2][i] = __float_as_uint(r.dz);
    pool[3][i] = __float_as_uint(r.invx); pool[4][i] = __float_as_uint(r.invy); pool[5][i] = __float_as_uint(r.invz);
    pool[6][i] = __float_as_uint(r.wx); pool[7][i] = __float_as_uint(r.wy); pool[8][i] = __float_as_uint(r.wz);
    pool[9][i] = (uint32_t)r.gx; pool[10][i] = (uint32_t)r.gy; pool[11][i] = (uint32_t)r.gz;
    pool[12][i] = (uint32_t)r.limit; pool[13][i] = r.flags;
    if (COUNT) { pool[14][i] = r.c0; pool[15][i] = r.c1; }
}

template <int COUNT>
__device__ __forceinline__ void pool_load(uint32_t (*pool)[kPoolRays], int i, PoolRay &r) {
    r.dx = __uint_as_float(pool[0][i]); r.dy = __uint_as_float(pool[1][i]); r.dz = __uint_as_float(pool[2][i]);
    r.invx = __uint_as_float(pool[3][i]); r.invy = __uint_as_float(pool[4][i]); r.invz = __uint_as_float(pool[5][i]);
    r.wx = __uint_as_float(pool[6][i]); r.wy = __uint_as_float(pool[7][i]); r.wz = __uint_as_float(pool[8][i]);
    r.gx = (int)pool[9][i]; r.gy = (int)pool[10][i]; r.gz = (int)pool[11][i];
    r.limit = (int)pool[12][i]; r.flags = pool[13][i];
    if (COUNT) { r.c0 = pool[14][i]; r.c1 = pool[15][i]; } else { r.c0 = r.c1 = 0; }
}

// Set up a pooled ray from an origin / direction exactly as traceMap does (map.glsl:85-102).
// Returns false when the ray must take the generic path (non-finite reciprocals, far origins).
__device__ __forceinline__ bool pool_init_ray(const WorldCompact &w, float ox, float oy, float oz, float dx, float dy, float dz,
                                              int max_steps, bool allow_seal, uint32_t pixel, PoolRay &r) {
    uint32_t zero = 0;
    if (dx == 0.0f) { dx = 0.001f; zero |= 1u; }
    if (dy == 0.0f) { dy = 0.001f; zero |= 2u; }
    if (dz == 0.0f) { dz = 0.001f; zero |= 4u; }
    r.dx = dx; r.dy = dy; r.dz = dz;
    r.invx = 1.0f / dx; r.invy = 1.0f / dy; r.invz = 1.0f / dz;
    const float o8x = ox * 8.0f, o8y = oy * 8.0f, o8z = oz * 8.0f;
    const bool sane = fabsf(r.invx) < 1e30f && fabsf(r.invy) < 1e30f && fabsf(r.invz) < 1e30f &&
                      fabsf(dx) < 1e30f && fabsf(dy) < 1e30f && fabsf(dz) < 1e30f &&
                      fabsf(o8x) < 1048576.0f && fabsf(o8y) < 1048576.0f && fabsf(o8z) < 1048576.0f && max_steps > 0;
    r.gx = __float2int_rz(o8x); r.gy = __float2int_rz(o8y); r.gz = __float2int_rz(o8z);
    r.wx = o8x - (float)r.gx; r.wy = o8y - (float)r.gy; r.wz = o8z - (float)r.gz;
    r.limit = 0;
    r.c0 = r.c1 = 0;
    // no map face in the direction of travel within reach of the trips (one block per trip at most)
    const int reach = max_steps + 2;
    const bool no_exit = (dx > 0.0f ? w.dim - 1 - (r.gx >> 3) : (r.gx >> 3)) >= reach && (dy > 0.0f ? w.dim - 1 - (r.gy >> 3) : (r.gy >> 3)) >= reach &&
                         (dz > 0.0f ? w.dim - 1 - (r.gz >> 3) : (r.gz >> 3)) >= reach;
    r.flags = kFlagMx | (zero << kFlagZeroShift) | ((allow_seal && dy > 0.0f && no_exit) ? kFlagClimbs : 0u) | (pixel << 16);
    return sane;
}

// Advance every ray of the warp from trip `trip` up to (not including the lookup of) trip `trip_end`.
// On return r.limit == kDead marks a finished ray whose result is in `out`.  All 32 lanes must call.
template <int COUNT>
__device__ __forceinline__ void pool_advance(const WorldCompact &w, const uint32_t *s_masks, const uint32_t *__restrict__ g_masks, PoolRay &r, Hit &out,
                                             int trip, const int trip_end, const int max_steps, const int bound) {
    const bool posx = r.dx > 0.0f, posy = r.dy > 0.0f, posz = r.dz > 0.0f;
    bool big = (r.flags & kFlagBig) != 0u;
    bool mx = (r.flags & kFlagMx) != 0u, my = (r.flags & kFlagMy) != 0u;
    const bool climbs = (r.flags & kFlagClimbs) != 0u;
    const float step = big ? 8.0f : 1.0f;
    float tgx = posx ? step : 0.0f, tgy = posy ? step : 0.0f, tgz = posz ? step : 0.0f;
    float rsx = posx ? 0.0f : step * 0.999f, rsy = posy ? 0.0f : step * 0.999f, rsz = posz ? 0.0f : step * 0.999f;
    const int istep = big ? 8 : 1;
    int isx = posx ? istep : -istep, isy = posy ? istep : -istep, isz = posz ? istep : -istep;
    int gx = r.gx, gy = r.gy, gz = r.gz;
    float wx = r.wx, wy = r.wy, wz = r.wz;
    const float dx = r.dx, dy = r.dy, dz = r.dz, invx = r.invx, invy = r.invy, invz = r.invz;
    int limit = r.limit;
    uint32_t cmat = 0;
    const uint32_t cd1 = w.cd1;

    // sealed-ray test (c) at the phase boundary, all candidate rays at once (see trace.cuh)
    if (COUNT != 1 && climbs && limit < kDead && (big || trip == 0) &&
        sky_sealed(w.clear4, w.clear64, w.dim, w.y_clear, ((float)gx + wx) * 0.125f, ((float)gy + wy) * 0.125f, ((float)gz + wz) * 0.125f, dx, dy, dz, max_steps - trip)) {
        out.trips = (uint32_t)max_steps;  // iteration-cap miss (map.glsl:167)
        limit = kDead;
    }

    while (trip < trip_end) {
        // ---- lookups (map.glsl:107-144): every live lane, see trace_map_fast ---------------
        if (limit < kDead) {
            if ((unsigned)gx >= (unsigned)bound || (unsigned)gy >= (unsigned)bound || (unsigned)gz >= (unsigned)bound) {
                out.exit_kind = 2;
                out.trips = (uint32_t)trip;
                limit = kDead;
            } else {
                if (COUNT == 2) r.c0++;  // lookups performed
                const uint32_t px = (uint32_t)gx + __float2uint_rz(wx);
                const uint32_t py = (uint32_t)gy + __float2uint_rz(wy);
                const uint32_t pz = (uint32_t)gz + __float2uint_rz(wz);
                uint32_t mat;
                int n_free = 0;
                if (!big && cmat != 0u && (((px ^ out.px) | (py ^ out.py) | (pz ^ out.pz)) < 8u)) {
                    mat = cmat;  // still inside the block of the previous lookup
                    if (COUNT == 1) r.c0++;
                } else {
                    const uint32_t e = __ldg(&w.chunks2[(px >> 6) + cd1 * ((py >> 6) + (pz >> 6) * cd1)]);
                    if ((int)e < 0) {
                        n_free = (int)(e & 0xFFu);
                        mat = 0u;
                    } else {
                        if (COUNT == 1 && e < w.n_real_bricks) r.c0++;
                        const uint32_t b8 = __ldg(&w.bricks8[e * 512u + (((px >> 3) & 7u) | (py & 0x38u) | ((pz & 0x38u) << 3))]);
                        const bool is_mat = b8 < kMatLimit;
                        n_free = is_mat ? 0 : (int)(b8 - kMatLimit);
                        mat = is_mat ? b8 : 0u;
                    }
                    cmat = mat;
                }
                out.px = px; out.py = py; out.pz = pz;
                if (COUNT == 1) n_free = 0;  // exact reference counters need every lookup
                limit = max(limit, trip + 1 + n_free);
                const bool seal = COUNT != 1 && mat == 0u && (limit >= max_steps || (climbs && (gy >> 3) >= w.y_clear));
                limit = min(limit, max_steps);
                if (seal) {  // nothing but empty in-map blocks until the iteration cap
                    out.trips = (uint32_t)max_steps;
                    limit = kDead;
                } else if (mat != 0u) {
                    if (COUNT == 1) r.c1++;
                    const uint32_t bit = (px & 7u) | ((py & 7u) << 3) | ((pz & 7u) << 6);
                    const uint32_t word = mat < (uint32_t)kPoolMaskMats ? s_masks[mat * 16u + (bit >> 5)] : __ldg(&g_masks[mat * 16u + (bit >> 5)]);
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

        // ---- k DDA steps for the whole warp, uniform trip count --------------------------
        int k = __reduce_min_sync(0xFFFFFFFFu, limit - trip);
        if (k >= kDead / 2) break;        // no live lane left in this warp
        k = min(k, trip_end - trip);      // stop at the phase boundary (>= 1 here)
        for (int j = 1; j < k; ++j) dda_step(gx, gy, gz, wx, wy, wz, isx, isy, isz, tgx, tgy, tgz, invx, invy, invz, dx, dy, dz, rsx, rsy, rsz);
        {
            int mxi, myi;
            dda_step_last(gx, gy, gz, wx, wy, wz, mxi, myi, isx, isy, isz, tgx, tgy, tgz, invx, invy, invz, dx, dy, dz, rsx, rsy, rsz);
            mx = mxi != 0;
            my = myi != 0;
        }
        trip += k;
        if (trip >= max_steps) {  // iteration cap: miss (map.glsl:167)
            if (limit < kDead) {
                out.trips = (uint32_t)trip;
                limit = kDead;
            }
            break;
        }
    }
    r.gx = gx; r.gy = gy; r.gz = gz;
    r.wx = wx; r.wy = wy; r.wz = wz;
    r.limit = limit;
    r.flags = (r.flags & ~(kFlagBig | kFlagMx | kFlagMy)) | (big ? kFlagBig : 0u) | (mx ? kFlagMx : 0u) | (my ? kFlagMy : 0u);
}

__device__ __forceinline__ void hit_reset(Hit &h) {
    h.data = 0;
    h.hx = h.hy = h.hz = -1.0f;
    h.px = h.py = h.pz = 0xFFFFFFFFu;
    h.block = 0;
    h.face = 0;
    h.exit_kind = 1;
    h.trips = 0;
}

// next phase boundary after trip b: 0 -> 4 -> 16 -> 32 -> 64 -> 128 -> ...
__device__ __forceinline__ int next_boundary(int b) { return b == 0 ? 4 : (b == 4 ? 16 : 2 * b); }

// Compact the rays with keep == true of this CTA pass into pool[][base...] in thread order.
// Returns nothing; s_base is advanced by the number of kept rays.  All 128 threads must call.
template <int COUNT>
__device__ __forceinline__ void pool_compact(uint32_t (*pool)[kPoolRays], int *s_cnt, int *s_base, bool keep, const PoolRay &r) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = __popc(m);
    __syncthreads();  // also orders every thread's pool_load of this pass before the stores below
    int base = *s_base;
    for (int i = 0; i < warp; ++i) base += s_cnt[i];
    if (keep) pool_store<COUNT>(pool, base + __popc(m & ((1u << lane) - 1u)), r);
    __syncthreads();
    if (threadIdx.x == 0) *s_base += s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
    // the next reader of *s_base / s_cnt is separated by the __syncthreads of the next pass
}

__device__ __forceinline__ void pool_stage_masks(uint32_t *s_masks, const uint32_t *__restrict__ g_masks, uint32_t n_mats) {
    const uint32_t n = min(n_mats + 1u, (uint32_t)kPoolMaskMats) * 16u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_masks[i] = __ldg(&g_masks[i]);
}

// pixel of pool entry / thread slot: 4 warps as 2x2 tiles of 8x4 pixels, two 16x8 halves per CTA tile
__device__ __forceinline__ uint32_t pool_pixel_id(int slot) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t x = (warp & 1u) * 8u + (lane & 7u);
    const uint32_t y = (uint32_t)slot * 8u + (warp >> 1) * 4u + (lane >> 3);
    return y * kPoolTile + x;
}

// ---- primary pass, pooled ---------------------------------------------------------------------
template <int COUNT, bool HITBUF, bool BATCH>
__global__ void __launch_bounds__(kThreads, UVT_MIN_BLOCKS) primary_pool_kernel(WorldArgs<WorldCompact> wa, const CamDev *__restrict__ cams, CamDev cam0,
                                                                                ViewDev v, GBufDev gb, DevCounters *counters) {
    __shared__ uint32_t s_pool[PoolWords<COUNT>::n][kPoolRays];
    __shared__ uint32_t s_masks[kPoolMaskMats * 16];
    __shared__ int s_cnt[4];
    __shared__ int s_base, s_count;
    WorldCompact w = wa.w;
    pool_stage_masks(s_masks, wa.masks, wa.n_mats);
    if (threadIdx.x == 0) { s_base = 0; s_count = kPoolRays; }
    __syncthreads();
    const CamDev &cam = BATCH ? cams[blockIdx.z] : cam0;
    const int max_steps = (int)v.max_steps, bound = (int)(8u * v.map_dim);
    uint32_t n_rays = 0, n_hits = 0, t_in = 0, t_chunk = 0, t_block = 0;

    int trip_b = 0;
    for (;;) {
        const int trip_e = min(next_boundary(trip_b), max_steps);
        const int n_in = s_count;
        for (int c = 0; c * kThreads < n_in; ++c) {
            const int i = c * kThreads + (int)threadIdx.x;
            PoolRay r;
            Hit h;
            hit_reset(h);
            bool generic = false;
            float odx = 0.0f, ody = 0.0f, odz = 1.0f, sx = 0.0f, sy = 0.0f, sz = 0.0f;  // generic path only
            if (trip_b == 0) {
                // phase 0: rays are born from the pixels of slot c
                const uint32_t pid = pool_pixel_id(c);
                const uint32_t x = blockIdx.x * kPoolTile + (pid & 15u), ly = blockIdx.y * kPoolTile + (pid >> 4);
                uint32_t y;
                const bool valid = v.global_row(ly, y) && x < v.W;
                r.limit = kDead; r.flags = pid << 16; r.c0 = r.c1 = 0;
                r.dx = r.dy = r.dz = r.invx = r.invy = r.invz = 1.0f; r.wx = r.wy = r.wz = 0.0f; r.gx = r.gy = r.gz = 0;
                if (valid) {
                    primary_ray(cam, v, x, y, odx, ody, odz, sx, sy, sz);
                    generic = !pool_init_ray(w, sx, sy, sz, odx, ody, odz, max_steps, COUNT != 1, pid, r);
                    if (generic) r.limit = kDead;
                    if (COUNT) n_rays++;
                }
                if (generic) {  // rare lanes: the generic loop is the specification for corner cases
                    TripCounts tc = {0, 0, 0};
                    w.smem_masks = wa.masks;  // the generic view reads the masks through a generic pointer
                    trace_map<WorldCompact, COUNT == 1>(w, sx, sy, sz, odx, ody, odz, max_steps, bound, h, tc);
                    r.c0 = tc.t_chunk; r.c1 = tc.t_block;
                }
            } else if (i < n_in) {
                pool_load<COUNT>(s_pool, i, r);
            } else {
                r.limit = kDead; r.flags = 0; r.c0 = r.c1 = 0;
                r.dx = r.dy = r.dz = r.invx = r.invy = r.invz = 1.0f; r.wx = r.wy = r.wz = 0.0f; r.gx = r.gy = r.gz = 0;
            }
            const bool had_ray = r.limit < kDead || generic;
            pool_advance<COUNT>(w, s_masks, wa.masks, r, h, trip_b, trip_e, max_steps, bound);  // all 32 lanes
            const bool alive = r.limit < kDead;
            if (had_ray && !alive) {
                // ---- the ray ended in this phase: write its pixel (primary.comp.glsl:58-68)
                const uint32_t pid = r.flags >> 16;
                const uint32_t x = blockIdx.x * kPoolTile + (pid & 15u), ly = blockIdx.y * kPoolTile + (pid >> 4);
                const size_t px = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
                float dist = -1.0f;
                if (h.data != 0) {
                    gb.albedo[px] = h.data;
                    gb.normal[px] = normal_rgba8(h.face);
                    gb.position[px] = make_float4(ceilf(h.hx) / 8.0f, ceilf(h.hy) / 8.0f, ceilf(h.hz) / 8.0f, 1.0f);
                    if (HITBUF) {
                        const float ex = h.hx / 8.0f - cam.pos[0], ey = h.hy / 8.0f - cam.pos[1], ez = h.hz / 8.0f - cam.pos[2];
                        dist = sqrtf(ex * ex + ey * ey + ez * ez);
                    }
                    if (COUNT) n_hits++;
                } else {
                    // sky colour from the UNPATCHED direction (the 0.001 patch lives inside traceMap only)
                    float rdx, rdy, rdz;
                    if (generic) { rdx = odx; rdy = ody; rdz = odz; }
                    else {
                        const uint32_t zero = (r.flags >> kFlagZeroShift) & 7u;
                        rdx = (zero & 1u) ? 0.0f : r.dx; rdy = (zero & 2u) ? 0.0f : r.dy; rdz = (zero & 4u) ? 0.0f : r.dz;
                    }
                    float cr, cg, cb;
                    sky_dome2(rdx, rdy, rdz, cr, cg, cb);
                    gb.albedo[px] = pack_rgba8(cr, cg, cb, 1.0f);
                    gb.normal[px] = 0xFFFFFFFFu;
                    gb.position[px] = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
                    h.px = h.py = h.pz = 0xFFFFFFFFu;
                }
                if (HITBUF) store_hit(gb.hit, px, h, dist);
                if (COUNT == 1) { t_in += h.trips; t_chunk += r.c0; t_block += r.c1; }
                if (COUNT == 2) t_in += r.c0;
            }
            pool_compact<COUNT>(s_pool, s_cnt, &s_base, alive, r);
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_count = s_base; s_base = 0; }
        __syncthreads();
        trip_b = trip_e;
        if (trip_b >= max_steps || s_count == 0) break;
    }
    if (COUNT) {
        warp_add(&counters->rays, n_rays);
        warp_add(&counters->t_in, t_in);
        warp_add(&counters->t_chunk, t_chunk);
        warp_add(&counters->t_block, t_block);
        warp_add(&counters->hits, n_hits);
    }
}

// ---- secondary (sun shadow) pass, pooled ---------------------------------------------------------
template <int COUNT>
__global__ void __launch_bounds__(kThreads, UVT_MIN_BLOCKS) secondary_pool_kernel(WorldArgs<WorldCompact> wa, ViewDev v, GBufDev gb, DevCounters *counters) {
    __shared__ uint32_t s_pool[PoolWords<COUNT>::n][kPoolRays];
    __shared__ uint32_t s_masks[kPoolMaskMats * 16];
    __shared__ int s_cnt[4];
    __shared__ int s_base, s_count;
    WorldCompact w = wa.w;
    pool_stage_masks(s_masks, wa.masks, wa.n_mats);
    if (threadIdx.x == 0) { s_base = 0; s_count = kPoolRays; }
    __syncthreads();
    const int max_steps = (int)v.max_steps, bound = (int)(8u * v.map_dim);
    uint32_t n_rays = 0, n_hits = 0, n_early = 0, t_in = 0, t_chunk = 0, t_block = 0;

    int trip_b = 0;
    for (;;) {
        const int trip_e = min(next_boundary(trip_b), max_steps);
        const int n_in = s_count;
        for (int c = 0; c * kThreads < n_in; ++c) {
            const int i = c * kThreads + (int)threadIdx.x;
            PoolRay r;
            Hit h;
            hit_reset(h);
            bool generic = false;
            float ox = 0.0f, oy = 0.0f, oz = 0.0f;
            r.limit = kDead; r.flags = 0; r.c0 = r.c1 = 0;
            r.dx = r.dy = r.dz = r.invx = r.invy = r.invz = 1.0f; r.wx = r.wy = r.wz = 0.0f; r.gx = r.gy = r.gz = 0;
            if (trip_b == 0) {
                const uint32_t pid = pool_pixel_id(c);
                const uint32_t x = blockIdx.x * kPoolTile + (pid & 15u), ly = blockIdx.y * kPoolTile + (pid >> 4);
                uint32_t y;
                const bool valid = v.global_row(ly, y) && x < v.W;
                r.flags = pid << 16;
                if (valid) {
                    const size_t px = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
                    const float4 pos = gb.position[px];
                    if (pos.x < 0.0f || pos.y < 0.0f || pos.z < 0.0f) {  // secondary.comp.glsl:26-29
                        gb.illum[px] = 0u;
                        if (COUNT) n_early++;
                    } else {
                        const uint32_t nrm = gb.normal[px];  // :36-37
                        ox = pos.x + ((float)(nrm & 255u) / 255.0f) * 0.001f;
                        oy = pos.y + ((float)((nrm >> 8) & 255u) / 255.0f) * 0.001f;
                        oz = pos.z + ((float)((nrm >> 16) & 255u) / 255.0f) * 0.001f;
                        generic = !pool_init_ray(w, ox, oy, oz, UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, max_steps, COUNT != 1, pid, r);
                        if (generic) r.limit = kDead;
                        if (COUNT) n_rays++;
                    }
                }
                if (generic) {
                    TripCounts tc = {0, 0, 0};
                    w.smem_masks = wa.masks;
                    trace_map<WorldCompact, COUNT == 1>(w, ox, oy, oz, UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, max_steps, bound, h, tc);
                    r.c0 = tc.t_chunk; r.c1 = tc.t_block;
                }
            } else if (i < n_in) {
                pool_load<COUNT>(s_pool, i, r);
            }
            const bool had_ray = r.limit < kDead || generic;
            pool_advance<COUNT>(w, s_masks, wa.masks, r, h, trip_b, trip_e, max_steps, bound);  // all 32 lanes
            const bool alive = r.limit < kDead;
            if (had_ray && !alive) {
                const uint32_t pid = r.flags >> 16;
                const uint32_t x = blockIdx.x * kPoolTile + (pid & 15u), ly = blockIdx.y * kPoolTile + (pid >> 4);
                const size_t px = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
                bool shadowed = h.data != 0;
                if (COUNT && shadowed) n_hits++;
                if (v.entities && !shadowed) {  // secondary.comp.glsl:42; an entity hit only matters when the terrain ray missed
                    const float4 pos = gb.position[px];
                    const uint32_t nrm = gb.normal[px];
                    const float sox = pos.x + ((float)(nrm & 255u) / 255.0f) * 0.001f;
                    const float soy = pos.y + ((float)((nrm >> 8) & 255u) / 255.0f) * 0.001f;
                    const float soz = pos.z + ((float)((nrm >> 16) & 255u) / 255.0f) * 0.001f;
                    const float ex = sox - h.hx / 8.0f, ey = soy - h.hy / 8.0f, ez = soz - h.hz / 8.0f;
                    shadowed = trace_entities(sox, soy, soz, UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, sqrtf(ex * ex + ey * ey + ez * ez));
                }
                gb.illum[px] = pack_rgba8(UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, shadowed ? -0.3f : 0.3f);
                if (COUNT == 1) { t_in += h.trips; t_chunk += r.c0; t_block += r.c1; }
                if (COUNT == 2) t_in += r.c0;
            }
            pool_compact<COUNT>(s_pool, s_cnt, &s_base, alive, r);
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_count = s_base; s_base = 0; }
        __syncthreads();
        trip_b = trip_e;
        if (trip_b >= max_steps || s_count == 0) break;
    }
    if (COUNT) {
        warp_add(&counters->rays, n_rays);
        warp_add(&counters->t_in, t_in);
        warp_add(&counters->t_chunk, t_chunk);
        warp_add(&counters->t_block, t_block);
        warp_add(&counters->hits, n_hits);
        warp_add(&counters->early_out, n_early);
    }
}

}  // namespace uvt
