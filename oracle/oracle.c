/*
 * oracle.c — CPU restatement of the reference's per-pixel voxel ray-traversal pass.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under the product package may import, link or execute
 * this file; it is the parity checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * The reference is GLSL 4.50 compute + Zig/OpenGL host code and cannot be run as a program in
 * this image (no Zig, no GL, no Mesa: SURVEY.md §8c), and it holds no golden vectors or
 * known-answer tests for this path.  What pins this restatement:
 * (1) oracle/_ref/libglslref.so — the reference's OWN shader text (assets/shaders/ *.glsl, read
 *     where it lies, translated by syntactic rewrites only and compiled for the CPU against a
 *     GLSL-in-C++ shim: oracle/glsl_ref/).  tests/test_glsl_reference.py holds this file to it
 *     bit for bit: every G-buffer image, the illumination image and the final frame, traceMap,
 *     traceEntities and SkyDome2 ray by ray, on the default world, small worlds and the 4x world;
 * (2) the hand-derived known answers of SURVEY App. A.7;
 * (3) an independent pure-Python restatement (oracle/pyref.py) written from the shader text;
 * (4) committed golden fixtures (tests/golden/: outputs of the reference text and of this file).
 * Still unpinned: what a real GL driver does where GLSL leaves room (NaN handling of min/max/
 * clamp, pow, rounding of normalize, UNORM conversion) — both sides follow the contract below —
 * and the third-party noise / .vox code behind the INPUTS of the path.
 *
 * Every function cites the reference lines it follows (paths relative to the reference
 * checkout).  Arithmetic rules: fp32 everywhere, every + - * / individually
 * rounded (compile with -ffp-contract=off, no fast-math), float->int conversions truncate
 * toward zero and saturate, NaN converts to 0, GLSL min/max/clamp are the spec's ternaries.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- float -> int rules ------------------------------------------ */
static inline int32_t f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
static inline uint32_t f2u(float f) {
    if (f != f) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)f;
}
/* GLSL 4.50 §8.3: min(x,y) = y < x ? y : x;  max(x,y) = x < y ? y : x */
static inline float gmin(float x, float y) { return y < x ? y : x; }
static inline float gmax(float x, float y) { return x < y ? y : x; }
static inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
static inline float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

/* RGBA8 UNORM store: clamp to [0,1], scale by 255, add 0.5, drop the fraction (the
 * float->UNORM rule GPUs implement; it is what makes 0.3 -> 77, SURVEY App. A.7(iv):
 * 0.3f*255 is exactly 76.5 in fp32, which round-half-even would send to 76). */
static inline uint32_t unorm8(float c) {
    if (c != c) return 0;
    c = gclamp(c, 0.0f, 1.0f);
    return (uint32_t)(c * 255.0f + 0.5f);
}
static inline uint32_t pack_rgba8(float r, float g, float b, float a) {
    return unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (unorm8(a) << 24);
}

/* ---- camera.glsl:8 ------------------------------------------------------------------- */
static const float SUN_DIR[3] = {7.52185881e-01f, 6.58950984e-01f, 7.52185881e-01f};

/* SkyDome2: assets/shaders/camera.glsl:11-19.  rgb only (alpha is the constant 1). */
void orc_sky_dome2(const float rd[3], float col[3]) {
    const float sl = sqrtf(SUN_DIR[0] * SUN_DIR[0] + SUN_DIR[1] * SUN_DIR[1] + SUN_DIR[2] * SUN_DIR[2]);
    const float rl = sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
    const float dot = (SUN_DIR[0] / sl) * (rd[0] / rl) + (SUN_DIR[1] / sl) * (rd[1] / rl) + (SUN_DIR[2] / sl) * (rd[2] / rl);
    const float sun = gclamp(dot, 0.0f, 1.2f);
    const float k = rd[1] * 0.2f;
    const float base[3] = {0.6f, 0.71f, 0.75f};
    const float tilt[3] = {1.0f, 0.5f, 1.0f};
    const float warm[3] = {1.0f, 0.6f, 0.1f};
    const float glare[3] = {0.2f, 0.08f, 0.04f};
    const float p8 = powf(sun, 8.0f), p3 = powf(sun, 3.0f);
    for (int i = 0; i < 3; ++i) {
        float c = base[i] - k * tilt[i] + 0.15f * 0.5f;
        c += 0.4f * warm[i] * p8;
        c += glare[i] * p3;
        col[i] = c;
    }
}

/* intersectAABB: assets/shaders/map.glsl:21-29 */
static void intersect_aabb(const float o[3], const float d[3], const float bmin[3], const float bmax[3], float *t_near, float *t_far) {
    float t1[3], t2[3];
    for (int k = 0; k < 3; ++k) {
        const float tmin = (bmin[k] - o[k]) / d[k];
        const float tmax = (bmax[k] - o[k]) / d[k];
        t1[k] = gmin(tmin, tmax);
        t2[k] = gmax(tmin, tmax);
    }
    *t_near = gmax(gmax(t1[0], t1[1]), t1[2]);
    *t_far = gmin(gmin(t2[0], t2[1]), t2[2]);
}

/* map_getChunkFlags: map.glsl:31-36 */
static inline uint32_t chunk_flags(const orc_world *w, int32_t cx, int32_t cy, int32_t cz) {
    const int32_t cd = (int32_t)(w->dim / 8);
    if (cx < 0 || cy < 0 || cz < 0 || cx >= cd || cy >= cd || cz >= cd) return 0;
    return w->chunks[(size_t)cx + (size_t)cd * ((size_t)cy + (size_t)cz * (size_t)cd)];
}

/* map_getSubVoxel: map.glsl:57-60 (imageLoad of the 256^3 rgba8 atlas, then packUnorm4x8 — an exact byte round trip) */
static inline uint32_t sub_voxel(const orc_world *w, uint32_t mdl, int32_t x, int32_t y, int32_t z) {
    const uint32_t ox = (mdl & 31u) * 8u, oy = ((mdl / 32u) & 31u) * 8u, oz = ((mdl / 1024u) & 31u) * 8u;
    return w->atlas[(size_t)(ox + (uint32_t)x) + 256u * ((size_t)(oy + (uint32_t)y) + 256u * (size_t)(oz + (uint32_t)z))];
}

/* traceMap: assets/shaders/map.glsl:83-168 */
void orc_trace_map(const orc_world *w, const float origin[3], const float dir_in[3], int max_steps, orc_hit *out) {
    float d[3] = {dir_in[0], dir_in[1], dir_in[2]};
    for (int k = 0; k < 3; ++k)
        if (d[k] == 0.0f) d[k] = 0.001f; /* :85-90 */

    const int32_t bound = (int32_t)(8u * w->dim); /* :92 */
    int32_t sgn[3], pos[3];
    float inv[3];
    for (int k = 0; k < 3; ++k) {
        sgn[k] = f2i(gsign(d[k]));     /* :94 */
        pos[k] = (1 + sgn[k]) >> 1;    /* :95 */
        inv[k] = 1.0f / d[k];          /* :96 */
    }
    int min_idx = 0; /* :98 */
    int32_t g[3];
    float wi[3];
    for (int k = 0; k < 3; ++k) {
        const float o8 = origin[k] * 8.0f;
        g[k] = f2i(o8);               /* :101 */
        wi[k] = o8 - (float)g[k];     /* :102 */
    }
    uint32_t step = 0; /* :104 */

    memset(out, 0, sizeof *out);
    out->hit_pos[0] = out->hit_pos[1] = out->hit_pos[2] = -1.0f; /* :167 */
    out->p[0] = out->p[1] = out->p[2] = 0xFFFFFFFFu;
    out->distance = -1.0f;
    out->exit_kind = 1;

    int trip;
    for (trip = 0; trip < max_steps; ++trip) {
        if (g[0] >= bound || g[1] >= bound || g[2] >= bound || g[0] < 0 || g[1] < 0 || g[2] < 0) { /* :107,164 */
            out->exit_kind = 2;
            break;
        }
        out->t_in++;
        uint32_t p[3];
        for (int k = 0; k < 3; ++k) p[k] = (uint32_t)g[k] + f2u(wi[k]); /* :108 */

        /* map_getVoxel(ivec3(pos) >> 3): map.glsl:38-47 */
        const int32_t bx = (int32_t)p[0] >> 3, by = (int32_t)p[1] >> 3, bz = (int32_t)p[2] >> 3;
        uint32_t block = 0;
        const uint32_t blk_idx = chunk_flags(w, bx >> 3, by >> 3, bz >> 3);
        if (blk_idx > 0) {
            out->t_chunk++;
            block = w->bricks[(size_t)(blk_idx - 1) * 512u + (size_t)(bx % 8) + (size_t)((bz % 8) * 8 + (by % 8)) * 8u];
        }

        if (block != 0) {
            out->t_block++;
            const uint32_t sub = sub_voxel(w, block & 0xFFFFFFFu, (int32_t)p[0] % 8, (int32_t)p[1] % 8, (int32_t)p[2] % 8); /* :117 */
            if (sub != 0) {
                uint32_t face = 0; /* :119-125 */
                if (min_idx == 0) face = (uint32_t)(-pos[0] + 2);
                if (min_idx == 1) face = (uint32_t)(-pos[1] + 4);
                if (min_idx == 2) face = (uint32_t)(-pos[2] + 6);
                static const float normals[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}}; /* :72-79 */
                out->data = sub;
                for (int k = 0; k < 3; ++k) {
                    out->hit_pos[k] = (float)g[k] + wi[k]; /* :127 */
                    out->normal[k] = normals[face - 1][k];
                    out->p[k] = p[k];
                }
                out->face = face;
                out->block = block;
                out->exit_kind = 0;
                out->trips = (uint32_t)trip + 1;
                return;
            } else if (step != 0) { /* :131-135 */
                for (int k = 0; k < 3; ++k) {
                    g[k] += f2i(wi[k]);
                    wi[k] = wi[k] - floorf(wi[k]); /* fract */
                }
                step = 0;
            }
        } else if (step != 3) { /* :140-144 */
            for (int k = 0; k < 3; ++k) {
                wi[k] += (float)(g[k] & 7);
                g[k] -= g[k] & 7;
            }
            step = 3;
        }

        /* dda stepping: :157-162 */
        float t[3];
        for (int k = 0; k < 3; ++k) t[k] = ((float)(pos[k] << step) - wi[k]) * inv[k];
        min_idx = t[0] < t[1] ? (t[0] < t[2] ? 0 : 2) : (t[1] < t[2] ? 1 : 2);
        g[min_idx] += sgn[min_idx] * (1 << step); /* int(raySign << stepSize) */
        const float tm = t[min_idx];
        for (int k = 0; k < 3; ++k) wi[k] += d[k] * tm;
        wi[min_idx] = (float)((1 - pos[min_idx]) << step) * 0.999f;
    }
    out->trips = (uint32_t)trip;
}

/* traceEntities, live part: assets/shaders/map.glsl:172-201.  Returns 1 when HitInfo.data != 0. */
int orc_trace_entities(const float o[3], const float d[3], float max_distance) {
    static const float positions[5][3] = {{256.f, 21.f, 256.f}, {251.f, 21.f, 259.f}, {253.f, 21.f, 256.f}, {251.f, 21.f, 256.f}, {257.f, 21.f, 261.f}};
    float prev_d = INFINITY;
    int id = -1;
    for (int i = 0; i < 5; ++i) {
        const float dx = o[0] - positions[i][0], dy = o[1] - positions[i][1], dz = o[2] - positions[i][2];
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        if (dist >= max_distance) continue;
        const float bmax[3] = {positions[i][0] + 1.0f, positions[i][1] + 1.0f, positions[i][2] + 1.0f};
        float tn, tf;
        intersect_aabb(o, d, positions[i], bmax, &tn, &tf);
        if (tf >= tn && prev_d >= tf) {
            id = i;
            prev_d = tf;
        }
    }
    if (id >= 0) {
        const float bmax[3] = {positions[id][0] + 1.0f, positions[id][1] + 1.0f, positions[id][2] + 1.0f};
        float tn, tf;
        intersect_aabb(o, d, positions[id], bmax, &tn, &tf);
        if (tf >= tn) return 1; /* :199-201; everything after the return is dead code */
    }
    return 0;
}

/* traceEntities in full: assets/shaders/map.glsl:172-248 with the early `return HitInfo(0xFFFFFFFF, ...)` of :199 removed
 * when e->mode == 1.  Generalised only where the text has literals: the entity list (positions[], :173-179),
 * the model edge `bounds` (:203; the box edge follows as size/8 blocks so a voxel keeps the world's sub-voxel size) and
 * the step cap (:214).  With the defaults (5 literal positions, size 8, cap 64, model = atlas texels [0,8)^3) this is the
 * text as written.  Note what the text does NOT do: no zero patch of rayDir (sign(0) = 0 makes that axis "negative"
 * with an infinite reciprocal), no EPSILON on `withinGridCoords`. */
void orc_trace_entities_ex(const orc_world *w, const orc_entities *e, float epsilon, const float o[3], const float d[3],
                           float max_distance, orc_hit *out) {
    memset(out, 0, sizeof *out);
    out->p[0] = out->p[1] = out->p[2] = 0xFFFFFFFFu;
    out->distance = -1.0f;
    out->exit_kind = 1;
    const int32_t S = (int32_t)e->size;
    const float edge = (float)e->size / 8.0f; /* 1.0 for the literal 8^3 model: positions[i] + vec3(1.) */
    float prev_d = INFINITY; /* :182 */
    int id = -1;
    for (uint32_t i = 0; i < e->n; ++i) { /* :186-196 */
        const float dx = o[0] - e->pos[i][0], dy = o[1] - e->pos[i][1], dz = o[2] - e->pos[i][2];
        if (sqrtf(dx * dx + dy * dy + dz * dz) >= max_distance) continue;
        const float bmax[3] = {e->pos[i][0] + edge, e->pos[i][1] + edge, e->pos[i][2] + edge};
        float tn, tf;
        intersect_aabb(o, d, e->pos[i], bmax, &tn, &tf);
        if (tf >= tn && prev_d >= tf) {
            id = (int)i;
            prev_d = tf;
        }
    }
    if (id < 0) return;
    const float *P = e->pos[id];
    const float bmax[3] = {P[0] + edge, P[1] + edge, P[2] + edge};
    float tn, tf;
    intersect_aabb(o, d, P, bmax, &tn, &tf); /* :198 */
    if (!(tf >= tn)) return;
    if (e->mode == 0) { /* :199-201 as it runs */
        out->data = 0xFFFFFFFFu;
        for (int k = 0; k < 3; ++k) out->hit_pos[k] = P[k];
        out->block = (uint32_t)id;
        out->exit_kind = 3;
        return;
    }
    /* :203-211 */
    const float t0 = gmax(tn, 0.0f);
    float ro[3], inv[3], wi[3];
    int32_t sgn[3], pos[3], g[3];
    for (int k = 0; k < 3; ++k) {
        ro[k] = o[k] + t0 * d[k];
        sgn[k] = f2i(gsign(d[k]));
        pos[k] = (1 + sgn[k]) >> 1;
        inv[k] = 1.0f / d[k];
        g[k] = f2i((ro[k] - epsilon - P[k]) * 8.0f); /* :211 */
        wi[k] = (ro[k] - P[k]) * 8.0f - (float)g[k];  /* :212 */
    }
    int min_idx = 0; /* :208 */
    uint32_t trip;
    for (trip = 0; trip < e->max_steps; ++trip) { /* :214 */
        if (g[0] >= S || g[1] >= S || g[2] >= S || g[0] < 0 || g[1] < 0 || g[2] < 0) { /* :215, :243 */
            out->exit_kind = 2;
            break;
        }
        uint32_t p[3];
        for (int k = 0; k < 3; ++k) p[k] = ((uint32_t)g[k] + f2u(wi[k])) & (uint32_t)(S - 1); /* :216, :218 `& 7` */
        const uint32_t block = e->model ? e->model[(size_t)p[0] + (size_t)S * ((size_t)p[1] + (size_t)S * (size_t)p[2])]
                                        : w->atlas[(size_t)p[0] + 256u * ((size_t)p[1] + 256u * (size_t)p[2])];
        if (block != 0) { /* :220-230 */
            uint32_t face = 0;
            if (min_idx == 0) face = (uint32_t)(-pos[0] + 2);
            if (min_idx == 1) face = (uint32_t)(-pos[1] + 4);
            if (min_idx == 2) face = (uint32_t)(-pos[2] + 6);
            static const float normals[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
            out->data = block;
            for (int k = 0; k < 3; ++k) {
                out->hit_pos[k] = P[k] + ((float)g[k] + wi[k]) / 8.0f; /* :230: world space */
                out->normal[k] = normals[face - 1][k];
                out->p[k] = p[k];
            }
            out->face = face;
            out->block = (uint32_t)id;
            out->exit_kind = 3;
            out->trips = trip + 1;
            return;
        }
        for (int k = 0; k < 3; ++k) { /* :232-235 */
            g[k] += f2i(wi[k]);
            wi[k] = wi[k] - floorf(wi[k]);
        }
        float t[3]; /* :238-243 */
        for (int k = 0; k < 3; ++k) t[k] = ((float)pos[k] - wi[k]) * inv[k];
        min_idx = t[0] < t[1] ? (t[0] < t[2] ? 0 : 2) : (t[1] < t[2] ? 1 : 2);
        g[min_idx] += sgn[min_idx];
        const float tm = t[min_idx];
        for (int k = 0; k < 3; ++k) wi[k] += d[k] * tm;
        wi[min_idx] = (float)(1 - pos[min_idx]) * 0.999f;
    }
    out->trips = trip;
}

/* Ray generation: assets/shaders/primary.comp.glsl:31-43.  tan_half_fov = tanf(fov / 2). */
void orc_primary_ray(const orc_camera *cam, float tan_half_fov, uint32_t W, uint32_t H, uint32_t px, uint32_t py,
                     uint32_t map_dim, float epsilon, float origin[3], float dir[3], float start[3]) {
    float ux = (float)px / (float)W * 2.0f - 1.0f; /* :31 */
    float uy = (float)py / (float)H * 2.0f - 1.0f;
    uy *= (float)H / (float)W;                      /* :32 */
    ux *= tan_half_fov;                             /* :36 */
    uy *= tan_half_fov;
    /* C_view * vec4(uv,1,1): GLSL column j = cam_mat row j */
    const float in[4] = {ux, uy, 1.0f, 1.0f};
    float v[4];
    for (int i = 0; i < 4; ++i) {
        float acc = cam->cam_mat[0 * 4 + i] * in[0];
        acc = acc + cam->cam_mat[1 * 4 + i] * in[1];
        acc = acc + cam->cam_mat[2 * 4 + i] * in[2];
        acc = acc + cam->cam_mat[3 * 4 + i] * in[3];
        v[i] = acc;
    }
    const float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]); /* 4-component normalise */
    for (int k = 0; k < 3; ++k) {
        dir[k] = v[k] / len;
        origin[k] = cam->cam_pos[k];
    }
    const float bmin[3] = {0.0f, 0.0f, 0.0f};
    const float D = (float)map_dim;
    const float bmax[3] = {D, D, D};
    float tn, tf;
    intersect_aabb(origin, dir, bmin, bmax, &tn, &tf); /* :42 */
    const float t0 = gmax(tn, 0.0f);
    for (int k = 0; k < 3; ++k) start[k] = origin[k] + dir[k] * t0 - epsilon; /* :43 */
}

static void add_counters(orc_counters *c, const orc_hit *h) {
    c->rays++;
    c->t_in += h->t_in;
    c->t_chunk += h->t_chunk;
    c->t_block += h->t_block;
    c->hits += (h->data != 0);
}

static void merge_counters(orc_counters *dst, const orc_counters *src) {
    dst->rays += src->rays;
    dst->t_in += src->t_in;
    dst->t_chunk += src->t_chunk;
    dst->t_block += src->t_block;
    dst->hits += src->hits;
    dst->early_out += src->early_out;
}

/* One pixel of primary.comp.glsl main (:23-69): G-buffer texels + the explicit hit record. */
static void primary_pixel(const orc_world *w, const orc_camera *cam, const orc_params *prm, float thf, uint32_t W, uint32_t H,
                          uint32_t px, uint32_t py, uint32_t *albedo, uint32_t *normal, float *position, orc_hit_rec *rec,
                          orc_counters *cnt) {
    float o[3], d[3], s[3];
    orc_primary_ray(cam, thf, W, H, px, py, prm->map_dim, prm->epsilon, o, d, s);
    orc_hit h;
    orc_trace_map(w, s, d, (int)prm->primary_max_steps, &h);
    add_counters(cnt, &h);
    if (prm->ent && prm->ent->mode == 1) { /* :45-54, commented out in the reference: the entity composite */
        const float ex = o[0] - h.hit_pos[0] / 8.0f, ey = o[1] - h.hit_pos[1] / 8.0f, ez = o[2] - h.hit_pos[2] / 8.0f;
        orc_hit eh;
        orc_trace_entities_ex(w, prm->ent, prm->epsilon, o, d, sqrtf(ex * ex + ey * ey + ez * ez) + prm->epsilon, &eh); /* :47 */
        if (eh.data != 0) {
            *albedo = eh.data;                                                        /* :50 */
            *normal = pack_rgba8(eh.normal[0], eh.normal[1], eh.normal[2], 1.0f);     /* :51 */
            for (int k = 0; k < 3; ++k) position[k] = eh.hit_pos[k];                  /* :52: world space, no ceil/8 */
            position[3] = 1.0f;
            if (rec) {
                const float dx = eh.hit_pos[0] - o[0], dy = eh.hit_pos[1] - o[1], dz = eh.hit_pos[2] - o[2];
                rec->px = eh.p[0]; rec->py = eh.p[1]; rec->pz = eh.p[2];
                rec->block = 0x80000000u | eh.block;
                rec->color = eh.data;
                rec->distance = sqrtf(dx * dx + dy * dy + dz * dz);
                rec->trips = (uint16_t)eh.trips;
                rec->face = (uint8_t)eh.face;
                rec->exit_kind = 3;
            }
            return;
        }
    }
    if (h.data != 0) { /* :58-62 */
        *albedo = h.data;
        *normal = pack_rgba8(h.normal[0], h.normal[1], h.normal[2], 1.0f);
        for (int k = 0; k < 3; ++k) position[k] = ceilf(h.hit_pos[k]) / 8.0f;
        position[3] = 1.0f;
        const float dx = h.hit_pos[0] / 8.0f - o[0], dy = h.hit_pos[1] / 8.0f - o[1], dz = h.hit_pos[2] / 8.0f - o[2];
        h.distance = sqrtf(dx * dx + dy * dy + dz * dz);
    } else { /* :63-68 */
        float sky[3];
        orc_sky_dome2(d, sky);
        *albedo = pack_rgba8(sky[0], sky[1], sky[2], 1.0f);
        *normal = 0xFFFFFFFFu;
        for (int k = 0; k < 4; ++k) position[k] = -1.0f;
    }
    if (rec) {
        rec->px = h.p[0]; rec->py = h.p[1]; rec->pz = h.p[2];
        rec->block = h.block;
        rec->color = h.data;
        rec->distance = h.distance;
        rec->trips = (uint16_t)h.trips;
        rec->face = (uint8_t)h.face;
        rec->exit_kind = (uint8_t)h.exit_kind;
    }
}

/* primary.comp.glsl main: :23-69.  Row 0 is the bottom image row. */
void orc_primary(const orc_world *w, const orc_camera *cam, const orc_params *prm, uint32_t W, uint32_t H,
                 uint32_t *albedo, uint32_t *normal, float *position, orc_hit_rec *hits, orc_counters *counters) {
    const float thf = tanf(cam->fov / 2.0f);
    orc_counters total;
    memset(&total, 0, sizeof total);
#pragma omp parallel
    {
        orc_counters local;
        memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 4)
        for (int64_t py = 0; py < (int64_t)H; ++py) {
            for (uint32_t px = 0; px < W; ++px) {
                const size_t i = (size_t)py * W + px;
                primary_pixel(w, cam, prm, thf, W, H, px, (uint32_t)py, &albedo[i], &normal[i], &position[4 * i], hits ? &hits[i] : NULL, &local);
            }
        }
#pragma omp critical
        merge_counters(&total, &local);
    }
    if (counters) *counters = total;
}

/* The same for a LIST of pixels of the W x H frame (sampled parity checks of frames too large to render whole on the
 * CPU in test time): outputs are indexed by sample. */
void orc_primary_pixels(const orc_world *w, const orc_camera *cam, const orc_params *prm, uint32_t W, uint32_t H, size_t n,
                        const uint32_t *xs, const uint32_t *ys, uint32_t *albedo, uint32_t *normal, float *position, orc_hit_rec *hits) {
    const float thf = tanf(cam->fov / 2.0f);
#pragma omp parallel
    {
        orc_counters local;
        memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < (int64_t)n; ++i)
            primary_pixel(w, cam, prm, thf, W, H, xs[i], ys[i], &albedo[i], &normal[i], &position[4 * i], hits ? &hits[i] : NULL, &local);
    }
}

/* secondary.comp.glsl main: :18-51 */
void orc_secondary(const orc_world *w, const orc_params *prm, uint32_t W, uint32_t H,
                   const uint32_t *normal, const float *position, uint32_t *illum, orc_counters *counters) {
    orc_counters total;
    memset(&total, 0, sizeof total);
#pragma omp parallel
    {
        orc_counters local;
        memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 4)
        for (int64_t py = 0; py < (int64_t)H; ++py) {
            for (uint32_t px = 0; px < W; ++px) {
                const size_t i = (size_t)py * W + px;
                const float *pos = &position[4 * i];
                if (pos[0] < 0.0f || pos[1] < 0.0f || pos[2] < 0.0f) { /* :26-29 */
                    illum[i] = 0;
                    local.early_out++;
                    continue;
                }
                const uint32_t n = normal[i]; /* :36: RGBA8 UNORM read back as c/255 */
                const float nf[3] = {(float)(n & 255u) / 255.0f, (float)((n >> 8) & 255u) / 255.0f, (float)((n >> 16) & 255u) / 255.0f};
                float o[3];
                for (int k = 0; k < 3; ++k) o[k] = pos[k] + nf[k] * 0.001f; /* :37 */
                orc_hit h;
                orc_trace_map(w, o, SUN_DIR, (int)prm->shadow_max_steps, &h); /* :41 */
                add_counters(&local, &h);
                int ent = 0;
                if (prm->entities) {
                    const float dx = o[0] - h.hit_pos[0] / 8.0f, dy = o[1] - h.hit_pos[1] / 8.0f, dz = o[2] - h.hit_pos[2] / 8.0f;
                    const float maxd = sqrtf(dx * dx + dy * dy + dz * dz);
                    if (prm->ent) {
                        orc_hit eh;
                        orc_trace_entities_ex(w, prm->ent, prm->epsilon, o, SUN_DIR, maxd, &eh);
                        ent = eh.data != 0;
                    } else {
                        ent = orc_trace_entities(o, SUN_DIR, maxd); /* :42 */
                    }
                }
                const float a = (ent || h.data != 0) ? -0.3f : 0.3f; /* :45-48 */
                illum[i] = pack_rgba8(SUN_DIR[0], SUN_DIR[1], SUN_DIR[2], a); /* :50 */
            }
        }
#pragma omp critical
        merge_counters(&total, &local);
    }
    if (counters) *counters = total;
}

/* blit.fragment.glsl main: :23-36 with the 1:1 texel mapping of blit.vertex.glsl:5-14 */
void orc_blit(uint32_t W, uint32_t H, const uint32_t *albedo, const uint32_t *normal, const float *position,
              const uint32_t *illum, uint32_t *frame) {
#pragma omp parallel for schedule(static)
    for (int64_t py = 0; py < (int64_t)H; ++py) {
        for (uint32_t px = 0; px < W; ++px) {
            const size_t i = (size_t)py * W + px;
            const float tx = ((float)px + 0.5f) / (float)W, ty = ((float)py + 0.5f) / (float)H;
            float color[4], il[4], nrm[3];
            for (int k = 0; k < 4; ++k) {
                color[k] = (float)((albedo[i] >> (8 * k)) & 255u) / 255.0f;
                il[k] = (float)((illum[i] >> (8 * k)) & 255u) / 255.0f;
            }
            for (int k = 0; k < 3; ++k) nrm[k] = (float)((normal[i] >> (8 * k)) & 255u) / 255.0f;
            float rd[3] = {il[0], il[1], il[2]};
            float sky[4] = {0, 0, 0, 1.0f};
            /* :31: SkyDome2(rayPos + normal*0.001, illumination.xyz); the origin argument is unused by SkyDome2.
             * normalize(0) would be NaN, but then illumination.a == 0 too; GLSL 0*NaN = NaN, so keep the multiply
             * only when alpha != 0 (early-out pixels have il = 0 and the product is defined as 0 here). */
            if (illum[i] != 0) orc_sky_dome2(rd, sky);
            for (int k = 0; k < 4; ++k) color[k] = color[k] + (illum[i] != 0 ? il[3] * sky[k] : 0.0f);
            (void)nrm; (void)position;
            const float cx = tx - 0.5f, cy = ty - 0.5f;
            if (sqrtf(cx * cx + cy * cy) <= 0.002f) { /* :33 crosshair: mix(color, (1,1,1,0.4), 0.5) */
                const float target[4] = {1.0f, 1.0f, 1.0f, 0.4f};
                for (int k = 0; k < 4; ++k) color[k] = color[k] * (1.0f - 0.5f) + target[k] * 0.5f;
            }
            const float vx = tx * (1.0f - tx), vy = ty * (1.0f - ty); /* :15-21 */
            const float grad = powf(vx * vy * 15.0f, 0.6f * 0.3f);
            frame[i] = pack_rgba8(grad * color[0], grad * color[1], grad * color[2], grad * color[3]);
        }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
