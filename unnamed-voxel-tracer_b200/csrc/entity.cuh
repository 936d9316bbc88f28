// entity.cuh — traceEntities in full (assets/shaders/map.glsl:172-248) and the primary-pass entity composite the
// reference keeps commented out (primary.comp.glsl:45-54): SURVEY §8 row f3, "entities done properly".
//
// The reference as it RUNS returns at map.glsl:199 (entity boxes intersected as lines, shadow pass only); that live
// behaviour is compiled into secondary_kernel (kernels.cuh: trace_entities) and stays the default.  The passes here
// run only when the host asked for more — uvt_set_entity_mode(UVT_ENTITY_MODELS) makes the code behind that return
// live (a DDA over the entity's voxel model) together with the composite, uvt_set_entities replaces the five literal
// positions — as two extra launches after the primary / secondary kernels, so that the traversal kernels keep their
// register budget and the default frame is unchanged.
//
// Generalised only where the text has literals: the entity list (positions[], :173-179), the model edge `bounds`
// (:203; 8, 16 or 32 voxels — chicken.vox is 32^3, game.zig:114 — the box edge follows as size/8 blocks so a model
// voxel keeps the world's sub-voxel size) and the step cap (:214).  With the defaults this is the text as written,
// including what the text does NOT do: no zero patch of rayDir (sign(0) = 0 makes that axis "negative" with an
// infinite reciprocal) and no EPSILON on withinGridCoords.
#pragma once

#include "kernels.cuh"

namespace uvt {

constexpr int kMaxEntities = 32;

struct EntityDev {
    uint32_t mode;        // 0 boxes (map.glsl:199 returns), 1 models
    uint32_t n;
    uint32_t size;        // model edge in voxels (8 / 16 / 32)
    uint32_t max_steps;   // 64 (map.glsl:214)
    const uint32_t *__restrict__ model;  // [size^3] RGBA8 texels, x + size * (y + size * z)
    float pos[kMaxEntities][3];
};

struct EntityHit {
    uint32_t data;     // HitInfo.data; 0 = nothing
    float hx, hy, hz;  // HitInfo.hit_pos, WORLD space (map.glsl:230)
    uint32_t face;
    uint32_t px, py, pz;  // model voxel
    uint32_t id;          // entity index
    uint32_t trips;
};

__device__ __forceinline__ int gsign_i(float x) { return x > 0.0f ? 1 : (x < 0.0f ? -1 : 0); }

__device__ inline void trace_entities_ex(const EntityDev &e, float epsilon, float ox, float oy, float oz, float dx, float dy, float dz,
                                         float max_distance, EntityHit &out) {
    out.data = 0;
    out.face = 0;
    out.trips = 0;
    const float edge = (float)e.size / 8.0f;
    float prev_d = __int_as_float(0x7f800000);  // :182
    int id = -1;
    const float dd = dx * dx + dy * dy + dz * dz;
    const bool finite_dir = dd < __int_as_float(0x7f800000);
    for (uint32_t i = 0; i < e.n; ++i) {  // :186-196
        const float Px = e.pos[i][0], Py = e.pos[i][1], Pz = e.pos[i][2];
        if (finite_dir) {
            // a line further than `edge` from the box centre misses the box (half diagonal 0.866 * edge): the exact slab test
            // below would find tFar < tNear by a margin far beyond its rounding.  Saves the six divisions for most boxes.
            const float h = 0.5f * edge;
            const float vx = Px + h - ox, vy = Py + h - oy, vz = Pz + h - oz;
            const float cx = vy * dz - vz * dy, cy = vz * dx - vx * dz, cz = vx * dy - vy * dx;
            if (cx * cx + cy * cy + cz * cz > dd * (edge * edge)) continue;
        }
        const float ex = ox - Px, ey = oy - Py, ez = oz - Pz;
        if (sqrtf(ex * ex + ey * ey + ez * ez) >= max_distance) continue;
        float tn, tf;
        intersect_aabb(ox, oy, oz, dx, dy, dz, Px, Py, Pz, Px + edge, Py + edge, Pz + edge, tn, tf);
        if (tf >= tn && prev_d >= tf) {
            id = (int)i;
            prev_d = tf;
        }
    }
    if (id < 0) return;
    const float Px = e.pos[id][0], Py = e.pos[id][1], Pz = e.pos[id][2];
    float tn, tf;
    intersect_aabb(ox, oy, oz, dx, dy, dz, Px, Py, Pz, Px + edge, Py + edge, Pz + edge, tn, tf);  // :198
    if (!(tf >= tn)) return;
    out.id = (uint32_t)id;
    if (e.mode == 0u) {  // :199-201 as it runs
        out.data = 0xFFFFFFFFu;
        out.hx = Px; out.hy = Py; out.hz = Pz;
        out.px = out.py = out.pz = 0xFFFFFFFFu;
        return;
    }
    // :203-212
    const int S = (int)e.size;
    const float t0 = gmax(tn, 0.0f);
    const float rx = ox + t0 * dx, ry = oy + t0 * dy, rz = oz + t0 * dz;
    const int sgx = gsign_i(dx), sgy = gsign_i(dy), sgz = gsign_i(dz);
    const int psx = (1 + sgx) >> 1, psy = (1 + sgy) >> 1, psz = (1 + sgz) >> 1;
    const float invx = 1.0f / dx, invy = 1.0f / dy, invz = 1.0f / dz;
    int gx = __float2int_rz((rx - epsilon - Px) * 8.0f), gy = __float2int_rz((ry - epsilon - Py) * 8.0f), gz = __float2int_rz((rz - epsilon - Pz) * 8.0f);
    float wx = (rx - Px) * 8.0f - (float)gx, wy = (ry - Py) * 8.0f - (float)gy, wz = (rz - Pz) * 8.0f - (float)gz;
    int mi = 0;  // :208
    uint32_t trip = 0;
    for (; trip < e.max_steps; ++trip) {  // :214
        if ((unsigned)gx >= (unsigned)S || (unsigned)gy >= (unsigned)S || (unsigned)gz >= (unsigned)S) break;  // :215, :243
        const uint32_t px = ((uint32_t)gx + __float2uint_rz(wx)) & (uint32_t)(S - 1);  // :216, :218
        const uint32_t py = ((uint32_t)gy + __float2uint_rz(wy)) & (uint32_t)(S - 1);
        const uint32_t pz = ((uint32_t)gz + __float2uint_rz(wz)) & (uint32_t)(S - 1);
        const uint32_t block = __ldg(&e.model[px + (uint32_t)S * (py + (uint32_t)S * pz)]);
        if (block != 0u) {  // :220-230
            out.data = block;
            out.face = mi == 0 ? (uint32_t)(2 - psx) : (mi == 1 ? (uint32_t)(4 - psy) : (uint32_t)(6 - psz));
            out.hx = Px + ((float)gx + wx) / 8.0f;
            out.hy = Py + ((float)gy + wy) / 8.0f;
            out.hz = Pz + ((float)gz + wz) / 8.0f;
            out.px = px; out.py = py; out.pz = pz;
            out.trips = trip + 1u;
            return;
        }
        gx += __float2int_rz(wx);  // :232-235
        gy += __float2int_rz(wy);
        gz += __float2int_rz(wz);
        wx = wx - floorf(wx);
        wy = wy - floorf(wy);
        wz = wz - floorf(wz);
        const float tx = ((float)psx - wx) * invx, ty = ((float)psy - wy) * invy, tz = ((float)psz - wz) * invz;  // :238
        mi = tx < ty ? (tx < tz ? 0 : 2) : (ty < tz ? 1 : 2);
        const float tm = mi == 0 ? tx : (mi == 1 ? ty : tz);
        wx += dx * tm;
        wy += dy * tm;
        wz += dz * tm;
        if (mi == 0) { gx += sgx; wx = (float)(1 - psx) * 0.999f; }
        else if (mi == 1) { gy += sgy; wy = (float)(1 - psy) * 0.999f; }
        else { gz += sgz; wz = (float)(1 - psz) * 0.999f; }
    }
    out.trips = trip;
}

// primary.comp.glsl:45-54 (commented out in the reference): composite the nearest entity over the terrain G-buffer.
// Runs after primary_kernel with the hit buffer on: hit.distance is distance(C_position, inter.hit_pos / 8) of :47.
template <bool BATCH>
__global__ void __launch_bounds__(256) entity_primary_kernel(EntityDev e, const CamDev *__restrict__ cams, CamDev cam0, ViewDev v, GBufDev gb) {
    const uint32_t x = blockIdx.x * 64u + (threadIdx.x & 63u);
    const uint32_t ly = v.row0 + blockIdx.y * 4u + (threadIdx.x >> 6);
    uint32_t y;
    if (!v.global_row(ly, y) || x >= v.W) return;
    const CamDev &cam = BATCH ? cams[blockIdx.z] : cam0;
    float dx, dy, dz, sx, sy, sz;
    primary_ray(cam, v, x, y, dx, dy, dz, sx, sy, sz);  // rayDir as main() holds it: traceMap patches zeros in its own copy only
    const size_t i = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
    uint32_t *rec = reinterpret_cast<uint32_t *>(gb.hit + i * 28u);
    float dist = __uint_as_float(rec[5]);
    if (rec[4] == 0u) {  // terrain miss: inter.hit_pos = vec3(-1) (map.glsl:167)
        const float ex = cam.pos[0] - (-1.0f / 8.0f), ey = cam.pos[1] - (-1.0f / 8.0f), ez = cam.pos[2] - (-1.0f / 8.0f);
        dist = sqrtf(ex * ex + ey * ey + ez * ez);
    }
    EntityHit h;
    trace_entities_ex(e, v.epsilon, cam.pos[0], cam.pos[1], cam.pos[2], dx, dy, dz, dist + v.epsilon, h);  // :47
    if (h.data == 0u) return;
    gb.albedo[i] = h.data;                                  // :50
    gb.normal[i] = normal_rgba8(h.face);                    // :51
    gb.position[i] = make_float4(h.hx, h.hy, h.hz, 1.0f);   // :52: world space as returned, no ceil / 8
    const float ex = h.hx - cam.pos[0], ey = h.hy - cam.pos[1], ez = h.hz - cam.pos[2];
    rec[0] = h.px; rec[1] = h.py; rec[2] = h.pz;
    rec[3] = 0x80000000u | h.id;
    rec[4] = h.data;
    rec[5] = __float_as_uint(sqrtf(ex * ex + ey * ey + ez * ez));
    rec[6] = (h.trips & 0xFFFFu) | ((h.face & 0xFFu) << 16) | (3u << 24);  // exit_kind 3: entity
}

// secondary.comp.glsl:42-48 for an entity set the compiled-in literal test does not cover: runs after secondary_kernel
// (launched with v.entities = 0) over the pixels the terrain left lit.
__global__ void __launch_bounds__(256) entity_shadow_kernel(EntityDev e, ViewDev v, GBufDev gb) {
    const uint32_t x = blockIdx.x * 64u + (threadIdx.x & 63u);
    const uint32_t ly = v.row0 + blockIdx.y * 4u + (threadIdx.x >> 6);
    uint32_t y;
    if (!v.global_row(ly, y) || x >= v.W) return;
    const size_t i = (size_t)blockIdx.z * gb.layer_pixels + (size_t)ly * v.W + x;
    const uint32_t il = gb.illum[i];
    if ((il >> 24) == 0u) return;  // early-out pixel (0) or already shadowed by the terrain (alpha 0): same texel either way
    const float4 pos = gb.position[i];
    const uint32_t normal = gb.normal[i];
    const float nx = unorm8_to_float(normal & 255u), ny = unorm8_to_float((normal >> 8) & 255u), nz = unorm8_to_float((normal >> 16) & 255u);
    const float ox = pos.x + nx * 0.001f, oy = pos.y + ny * 0.001f, oz = pos.z + nz * 0.001f;  // :36-37
    // the terrain ray missed: inter.hit_pos = vec3(-1), so maxDistance = distance(rayOrigin, vec3(-1) / 8) (:42)
    const float ex = ox - (-1.0f / 8.0f), ey = oy - (-1.0f / 8.0f), ez = oz - (-1.0f / 8.0f);
    EntityHit h;
    trace_entities_ex(e, v.epsilon, ox, oy, oz, UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, sqrtf(ex * ex + ey * ey + ez * ez), h);
    if (h.data != 0u) gb.illum[i] = pack_rgba8(UVT_SUN_X, UVT_SUN_Y, UVT_SUN_Z, -0.3f);  // :45-50
}

}  // namespace uvt
