// uvt.cu — the C ABI of include/uvt.h: context, device memory, uploads, repack, dispatch.
//
// One ctx = one CUDA device + one in-order stream (the reference is one GL context with one
// in-order queue: src/engine/graphics/shader.zig:113-117).  The host owns world / atlas /
// camera source data; the ctx owns all device memory and the pinned staging it hands out.
// There is NO CPU fallback anywhere in this file: every pass is a kernel launch.
#include "uvt.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the library is bound at run time (dlopen), see NcclApi

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"
#include "pool.cuh"
#include "procgen.cuh"
#include "entity.cuh"

using namespace uvt;

namespace {
thread_local std::string g_create_error;
}

struct uvt_ctx {
    uvt_params params;
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string error;
    uint64_t launches = 0;

    // ---- world staging (pinned host) and device copies
    uint32_t dim = 0, cd = 0;
    uint32_t *h_chunks = nullptr;
    uint32_t *h_bricks = nullptr;
    size_t h_capacity = 0;  // bricks
    bool staging_borrowed = false;  // h_chunks / h_bricks belong to the caller (uvt_world_use_staging)
    size_t n_bricks = 0;    // committed
    uint32_t *d_chunks = nullptr;
    uint32_t *d_bricks = nullptr;
    size_t d_brick_capacity = 0;
    uint8_t *d_bricks8 = nullptr;
    size_t d_brick8_capacity = 0;
    uint32_t *d_chunks2 = nullptr;  // fast-path chunk table [(cd+1)^3]
    size_t n_total_bricks = 0;      // brick slots in use in d_bricks8: real [0, n_bricks), spare, virtual [v_base, v_base + n_virtual)
    size_t v_base = 0;              // first virtual (clearance-only) brick slot; the spare slots below it take new real bricks
    size_t n_virtual = 0;
    uint8_t *d_rowmask = nullptr;       // [d_brick8_capacity][64] x-occupancy bits of every brick row
    uint32_t *d_brick_chunk = nullptr;  // [d_brick8_capacity] brick slot -> linear chunk index (kNoChunk: unused)
    unsigned int *d_tops32 = nullptr;   // [dim^2] column tops (block y + 1)
    uint8_t *d_field = nullptr;         // [(dim/8)^3] chunk distance field of the committed world
    uint8_t *d_field_tmp[2] = {nullptr, nullptr};  // separable-pass scratch, same size
    uint32_t *d_scratch = nullptr;      // staging of uvt_world_commit_region (kScratchWords)
    bool incremental_ok = false;        // the last full commit left everything uvt_world_commit_region needs
    int32_t y_clear = 0;            // max occupied block y + 1 (every block at or above is empty)
    uint16_t *d_clear4 = nullptr;   // [(dim/4)^2] dilated column-group tops for sky_sealed()
    uint16_t *d_sun1 = nullptr;     // [dim^2] sun clearance of the shadow pass per block column (sun_clear_kernel), built for sun_steps trips
    uint16_t *d_top3 = nullptr;     // [dim^2] column tops grown by [0, +1] columns (top2 of sun_clear_kernel): its input
    uint32_t sun_steps = 0;
    uint16_t *d_clear16 = nullptr;  // [ceil(dim/16)^2] their maxima over 16x16-block groups
    uint16_t *d_clear64 = nullptr;  // [ceil(dim/64)^2] ... over 64x64-block groups
    uint8_t *d_dense = nullptr;     // [dim^3] dense block grid (nullptr: not built — too large or disabled)
    bool dense_valid = false;
    bool world_committed = false;

    // ---- atlas: host copy by slot (slot = x/8 + 32*(y/8) + 1024*(z/8)), device [n_slots][512]
    std::vector<uint32_t> h_models;
    uint32_t n_slots = 0;
    uint32_t *d_models = nullptr;
    uint32_t d_model_slots = 0;
    bool atlas_dirty = true;

    // ---- materials (compact layout): id -> block word / colours / occupancy mask
    std::vector<uint32_t> mat_words;  // index = material id, [0] unused
    bool materials_dirty = true;
    bool compact_ok = false;
    uint32_t *d_mat_word = nullptr;   // [256]
    uint32_t *d_mat_color = nullptr;  // [256][512]
    uint32_t *d_mat_mask = nullptr;   // [256][16]

    // ---- cameras
    std::vector<uvt_camera> cams;
    CamDev cam0;
    CamDev *d_cams = nullptr;
    int d_cams_capacity = 0;
    bool have_camera = false;

    // ---- G-buffer
    uint32_t W = 0, H = 0, layers = 1;
    uint32_t band_rows = 8, n_parts = 1, part = 0, local_rows = 0;
    uint32_t *d_albedo = nullptr, *d_normal = nullptr, *d_illum = nullptr, *d_frame = nullptr;
    float4 *d_position = nullptr;
    uint8_t *d_hit = nullptr;
    size_t gbuf_pixels = 0;  // allocated pixels (all layers)
    uint32_t *frame_target = nullptr;
    bool frame_target_global_rows = false;
    void *shared_frame = nullptr;  // owned full-frame allocation exported over CUDA IPC (presenting rank)

    // ---- pipelined readback: two device snapshots + a copy stream
    cudaStream_t copy_stream = nullptr;
    void *snap[2] = {nullptr, nullptr};
    size_t snap_bytes[2] = {0, 0};
    cudaEvent_t snap_ready[2] = {}, snap_done[2] = {};
    bool snap_busy[2] = {false, false};
    int snap_next = 0;

    // ---- counters / timing
    DevCounters *d_counters = nullptr;
    uint8_t *d_pick = nullptr;
    bool timing = false;
    cudaEvent_t ev[4][2] = {};
    bool ev_valid[4] = {false, false, false, false};
    uint32_t *d_sink = nullptr;
    std::vector<void *> pinned;  // live uvt_alloc_pinned allocations

    // ---- device procgen (uvt_world_procgen_plan / _fill): temporaries kept between the two calls
    struct {
        uint16_t *vh = nullptr;
        uint32_t *seeds = nullptr, *deco = nullptr;
        uvt::pg::Tree *trees = nullptr;
        uint32_t n_trees = 0;
        size_t n_bricks = 0;
        bool planned = false;
    } pgen;

    // ---- frame in row chunks on two streams (uvt_set_frame_chunks): the tail of one chunk's pass runs under the next chunk's work
    uint32_t frame_chunks = 1;
    cudaStream_t side_stream = nullptr;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;

    // ---- entities (uvt_set_entity_mode / uvt_set_entities / uvt_entity_model_upload)
    uint32_t ent_mode = UVT_ENTITY_BOXES;
    std::vector<float> ent_pos;          // xyz per entity; empty: the five literal positions of map.glsl:173-179
    uint32_t ent_size = 8, ent_steps = 64;
    uint32_t *d_ent_model = nullptr;     // [ent_size^3]; nullptr: texels [0,8)^3 of the atlas (map.glsl:218)

    // ---- NCCL band exchange (uvt_nccl_init / uvt_dispatch_frame_nccl)
    ncclComm_t nccl = nullptr;
    int nccl_ranks = 0, nccl_rank = -1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t group_done[8] = {};   // band group g shaded (compute stream -> exchange stream)
    cudaEvent_t exchange_done = nullptr;
};

struct uvt_pipeline {
    uvt_ctx *ctx;
    uvt_pipeline_kind kind;
};

namespace {

int set_error(uvt_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    else g_create_error = buf;
    return code;
}

#define UVT_CUDA(ctx, expr)                                                                                     \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return set_error((ctx), e_ == cudaErrorMemoryAllocation ? UVT_ERR_OOM : UVT_ERR_CUDA, "%s: %s (%s:%d)", \
                             #expr, cudaGetErrorString(e_), __FILE__, __LINE__);                                \
    } while (0)

// several ctxs on different devices may live in one thread (uvt_group): every entry point selects its device
#define UVT_ENTER(ctx)                                  \
    do {                                                \
        int cur_ = -1;                                  \
        if (cudaGetDevice(&cur_) != cudaSuccess || cur_ != (ctx)->device) cudaSetDevice((ctx)->device); \
    } while (0)

#define UVT_REQUIRE(ctx, cond, msg)                                  \
    do {                                                             \
        if (!(cond)) return set_error((ctx), UVT_ERR_INVALID, "%s", (msg)); \
    } while (0)

uint32_t compute_local_rows(uint32_t H, uint32_t band_rows, uint32_t n_parts, uint32_t part) {
    if (n_parts == 1) return H;
    const uint32_t n_bands = (H + band_rows - 1) / band_rows;
    // bands part, part + n_parts, ...; all but possibly the globally last band are full
    uint32_t rows = 0;
    for (uint32_t b = part; b < n_bands; b += n_parts) rows += std::min(band_rows, H - b * band_rows);
    return rows;
}

// local row count including the padding of a ragged last band (storage is band-granular)
uint32_t storage_rows(uint32_t H, uint32_t band_rows, uint32_t n_parts, uint32_t part) {
    if (n_parts == 1) return H;
    const uint32_t n_bands = (H + band_rows - 1) / band_rows;
    uint32_t bands = 0;
    for (uint32_t b = part; b < n_bands; b += n_parts) ++bands;
    return bands * band_rows;
}

void free_gbuffer(uvt_ctx *c) {
    cudaFree(c->d_albedo); cudaFree(c->d_normal); cudaFree(c->d_position); cudaFree(c->d_illum);
    cudaFree(c->d_frame); cudaFree(c->d_hit);
    c->d_albedo = c->d_normal = c->d_illum = c->d_frame = nullptr;
    c->d_position = nullptr;
    c->d_hit = nullptr;
    c->gbuf_pixels = 0;
}

int alloc_gbuffer(uvt_ctx *c) {
    // GBuffer.resize recreates all four textures (gbuffer.zig:18-30)
    free_gbuffer(c);
    c->local_rows = compute_local_rows(c->H, c->band_rows, c->n_parts, c->part);
    const size_t rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
    const size_t px = (size_t)c->W * rows * c->layers;
    if (px == 0) return UVT_OK;
    UVT_CUDA(c, cudaMalloc(&c->d_albedo, px * 4));
    UVT_CUDA(c, cudaMalloc(&c->d_normal, px * 4));
    UVT_CUDA(c, cudaMalloc(&c->d_position, px * 16));
    UVT_CUDA(c, cudaMalloc(&c->d_illum, px * 4));
    UVT_CUDA(c, cudaMalloc(&c->d_frame, px * 4));
    if (c->params.flags & UVT_FLAG_HIT_BUFFER) UVT_CUDA(c, cudaMalloc(&c->d_hit, px * sizeof(uvt_hit)));
    // GL zero-initialises fresh storage
    UVT_CUDA(c, cudaMemsetAsync(c->d_albedo, 0, px * 4, c->stream));
    UVT_CUDA(c, cudaMemsetAsync(c->d_normal, 0, px * 4, c->stream));
    UVT_CUDA(c, cudaMemsetAsync(c->d_position, 0, px * 16, c->stream));
    UVT_CUDA(c, cudaMemsetAsync(c->d_illum, 0, px * 4, c->stream));
    UVT_CUDA(c, cudaMemsetAsync(c->d_frame, 0, px * 4, c->stream));
    c->gbuf_pixels = px;
    return UVT_OK;
}

size_t layer_pixels(const uvt_ctx *c) { return (size_t)c->W * storage_rows(c->H, c->band_rows, c->n_parts, c->part); }

void derive_cam(const uvt_camera &in, CamDev &out) {
    out.pos[0] = in.cam_pos[0]; out.pos[1] = in.cam_pos[1]; out.pos[2] = in.cam_pos[2];
    out.tan_half_fov = tanf(in.fov / 2.0f);  // primary.comp.glsl:36, evaluated once per frame on the host
    std::memcpy(out.mat, in.cam_mat, sizeof out.mat);
}

// Upload the atlas slots and (compact layout) rebuild material tables.
int ensure_ready(uvt_ctx *c) {
    UVT_REQUIRE(c, c->world_committed, "no world committed (uvt_world_alloc + uvt_world_commit first)");
    if (c->atlas_dirty) {
        const uint32_t slots = std::max<uint32_t>(c->n_slots, 1u);
        if (slots > c->d_model_slots) {
            cudaFree(c->d_models);
            c->d_models = nullptr;
            UVT_CUDA(c, cudaMalloc(&c->d_models, (size_t)slots * 512 * 4));
            c->d_model_slots = slots;
        }
        if (c->n_slots)
            UVT_CUDA(c, cudaMemcpyAsync(c->d_models, c->h_models.data(), (size_t)c->n_slots * 512 * 4, cudaMemcpyHostToDevice, c->stream));
        else
            UVT_CUDA(c, cudaMemsetAsync(c->d_models, 0, 512 * 4, c->stream));
        c->atlas_dirty = false;
        c->materials_dirty = true;
    }
    if (c->materials_dirty && c->compact_ok) {
        std::vector<uint32_t> words(256, 0), colors(256 * 512, 0), masks(256 * 16, 0);
        for (size_t m = 1; m < c->mat_words.size(); ++m) {
            words[m] = c->mat_words[m];
            const uint32_t slot = c->mat_words[m] & 32767u;
            if (slot >= c->n_slots) continue;  // unloaded model: reads as empty
            const uint32_t *tex = &c->h_models[(size_t)slot * 512];
            for (uint32_t b = 0; b < 512; ++b) {
                colors[m * 512 + b] = tex[b];
                if (tex[b] != 0) masks[m * 16 + (b >> 5)] |= 1u << (b & 31u);
            }
        }
        UVT_CUDA(c, cudaMemcpyAsync(c->d_mat_word, words.data(), 256 * 4, cudaMemcpyHostToDevice, c->stream));
        UVT_CUDA(c, cudaMemcpyAsync(c->d_mat_color, colors.data(), 256 * 512 * 4, cudaMemcpyHostToDevice, c->stream));
        UVT_CUDA(c, cudaMemcpyAsync(c->d_mat_mask, masks.data(), 256 * 16 * 4, cudaMemcpyHostToDevice, c->stream));
        UVT_CUDA(c, cudaStreamSynchronize(c->stream));  // the staging vectors die at scope exit
        c->materials_dirty = false;
    }
    return UVT_OK;
}

bool use_compact(const uvt_ctx *c) { return c->params.layout == UVT_LAYOUT_COMPACT && c->compact_ok; }

WorldArgs<WorldRef> world_ref(const uvt_ctx *c) {
    WorldArgs<WorldRef> a;
    a.w.chunks = c->d_chunks;
    a.w.bricks = c->d_bricks;
    a.w.models = c->d_models;
    a.w.cd = c->cd;
    a.w.n_slots = c->n_slots;
    a.masks = nullptr;
    a.n_mats = 0;
    return a;
}

WorldArgs<WorldCompact> world_compact(const uvt_ctx *c) {
    WorldArgs<WorldCompact> a;
    a.w.chunks = c->d_chunks;
    a.w.chunks2 = c->d_chunks2;
    a.w.cd1 = c->cd + 1;
    a.w.n_real_bricks = (uint32_t)c->n_bricks;
    a.w.y_clear = c->y_clear;
    a.w.dim = (int32_t)c->dim;
    a.w.clear4 = c->d_clear4;
    a.w.clear16 = c->d_clear16;
    a.w.sun1 = c->d_sun1;
    // the ray climbs at most (steps + 5) * u / (2 s + u) blocks; two blocks of slack under the top face
    a.w.sun_row_max = (int32_t)c->dim - 3 - (int32_t)std::ceil((double)(c->sun_steps + 5) * UVT_SUN_Y / (2.0 * UVT_SUN_X + UVT_SUN_Y));
    a.w.clear64 = c->d_clear64;
    a.w.dense = c->d_dense;
    a.w.bricks8 = c->d_bricks8;
    a.w.mat_word = c->d_mat_word;
    a.w.mat_color = c->d_mat_color;
    a.w.smem_masks = nullptr;
    a.w.g_masks = c->d_mat_mask;
    a.w.cd = c->cd;
    a.masks = c->d_mat_mask;
    a.n_mats = (uint32_t)(c->mat_words.size() - 1);
    return a;
}

bool ent_custom(const uvt_ctx *c) { return c->ent_mode == UVT_ENTITY_MODELS || !c->ent_pos.empty(); }

EntityDev make_entities(const uvt_ctx *c) {
    static const float literal[5][3] = {{256.f, 21.f, 256.f}, {251.f, 21.f, 259.f}, {253.f, 21.f, 256.f}, {251.f, 21.f, 256.f}, {257.f, 21.f, 261.f}};
    EntityDev e;
    e.mode = c->ent_mode;
    e.size = c->d_ent_model ? c->ent_size : 8u;
    e.max_steps = c->ent_steps;
    e.model = c->d_ent_model ? c->d_ent_model : c->d_models;  // atlas slot 0 = texels [0,8)^3, same x + 8 * (y + 8 * z) order
    if (c->ent_pos.empty()) {
        e.n = 5;
        std::memcpy(e.pos, literal, sizeof literal);
    } else {
        e.n = (uint32_t)(c->ent_pos.size() / 3);
        std::memcpy(e.pos, c->ent_pos.data(), c->ent_pos.size() * sizeof(float));
    }
    return e;
}

ViewDev make_view(const uvt_ctx *c, uint32_t max_steps) {
    ViewDev v;
    v.W = c->W; v.H = c->H;
    v.local_rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
    v.row0 = 0;
    v.band_rows = c->band_rows; v.n_parts = c->n_parts; v.part = c->part;
    v.map_dim = c->dim;
    v.max_steps = max_steps;
    v.epsilon = c->params.epsilon;
    // the compiled-in test covers the reference as it runs (five literal boxes, map.glsl:199 returns); anything else
    // is handled by the passes of entity.cuh after the traversal kernels
    v.entities = ((c->params.flags & UVT_FLAG_ENTITIES) && !ent_custom(c)) ? 1u : 0u;
    return v;
}

GBufDev make_gbuf(const uvt_ctx *c) {
    GBufDev g;
    g.albedo = c->d_albedo; g.normal = c->d_normal; g.position = c->d_position; g.illum = c->d_illum;
    g.frame = c->d_frame; g.hit = c->d_hit;
    g.layer_pixels = layer_pixels(c);
    return g;
}

FrameTarget make_target(const uvt_ctx *c) {
    FrameTarget t;
    t.ptr = c->frame_target ? c->frame_target : c->d_frame;
    t.global_rows = (c->frame_target && c->frame_target_global_rows) ? 1u : 0u;
    return t;
}

dim3 trace_grid(const uvt_ctx *c) {
    const uint32_t rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
    return dim3((c->W + kTileW - 1) / kTileW, (rows + kTileH - 1) / kTileH, c->layers);
}

struct PassTimer {
    uvt_ctx *c;
    int which;
    PassTimer(uvt_ctx *ctx, int w) : c(ctx), which(w) {
        if (c->timing) cudaEventRecord(c->ev[which][0], c->stream);
    }
    ~PassTimer() {
        if (c->timing) {
            cudaEventRecord(c->ev[which][1], c->stream);
            c->ev_valid[which] = true;
        }
    }
};

int check_launch(uvt_ctx *c, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(c, UVT_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    c->launches++;
    return UVT_OK;
}

int launch_rows(uvt_ctx *c, uint32_t row0, uint32_t nrows);  // the three passes over a row range (defined with the band exchange)

int pre_dispatch(uvt_ctx *c) {
    UVT_REQUIRE(c, c->W && c->H, "no G-buffer (uvt_resize first)");
    UVT_REQUIRE(c, c->have_camera, "no camera (uvt_set_camera first)");
    return ensure_ready(c);
}

template <class World, int COUNT>
void launch_primary_world(uvt_ctx *c, const WorldArgs<World> &wa, const ViewDev &v, const GBufDev &g, dim3 grid) {
    const bool hb = c->d_hit != nullptr, batch = c->layers > 1;
    const CamDev *cams = c->d_cams;
#define UVT_LAUNCH(HB, BATCH) primary_kernel<World, COUNT, HB, BATCH><<<grid, kTileThreads, 0, c->stream>>>(wa, cams, c->cam0, v, g, c->d_counters)
    if (hb) { if (batch) UVT_LAUNCH(true, true); else UVT_LAUNCH(true, false); }
    else { if (batch) UVT_LAUNCH(false, true); else UVT_LAUNCH(false, false); }
#undef UVT_LAUNCH
}

bool use_dense(const uvt_ctx *c) { return use_compact(c) && c->dense_valid && !(c->params.flags & UVT_FLAG_NO_DENSE); }

WorldArgs<WorldDense> world_dense(const uvt_ctx *c) {
    const WorldArgs<WorldCompact> s = world_compact(c);
    WorldArgs<WorldDense> a;
    static_cast<WorldCompact &>(a.w) = s.w;
    a.masks = s.masks;
    a.n_mats = s.n_mats;
    return a;
}

bool use_pool(const uvt_ctx *c) { return use_compact(c) && c->params.scheduler == UVT_SCHED_POOL; }

dim3 pool_grid(const uvt_ctx *c) {
    const uint32_t rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
    return dim3((c->W + kPoolTile - 1) / kPoolTile, (rows + kPoolTile - 1) / kPoolTile, c->layers);
}

template <int COUNT>
void launch_primary_pool(uvt_ctx *c, const ViewDev &v, const GBufDev &g) {
    const bool hb = c->d_hit != nullptr, batch = c->layers > 1;
    const dim3 grid = pool_grid(c);
    const auto wa = world_compact(c);
#define UVT_LAUNCH(HB, BATCH) primary_pool_kernel<COUNT, HB, BATCH><<<grid, kThreads, 0, c->stream>>>(wa, c->d_cams, c->cam0, v, g, c->d_counters)
    if (hb) { if (batch) UVT_LAUNCH(true, true); else UVT_LAUNCH(true, false); }
    else { if (batch) UVT_LAUNCH(false, true); else UVT_LAUNCH(false, false); }
#undef UVT_LAUNCH
}

// primary.comp.glsl:45-54 made live (uvt_set_entity_mode(UVT_ENTITY_MODELS)): the composite pass over `nrows` local rows from v.row0
int launch_entity_primary(uvt_ctx *c, const ViewDev &v, const GBufDev &g, uint32_t nrows, uint32_t layers) {
    if (c->ent_mode != UVT_ENTITY_MODELS || !(c->params.flags & UVT_FLAG_ENTITIES)) return UVT_OK;
    UVT_REQUIRE(c, c->d_hit, "entity models need the hit buffer (uvt_set_entity_mode allocates it; uvt_resize after it)");
    const dim3 grid((c->W + 63) / 64, (nrows + 3) / 4, layers);
    if (layers > 1) entity_primary_kernel<true><<<grid, 256, 0, c->stream>>>(make_entities(c), c->d_cams, c->cam0, v, g);
    else entity_primary_kernel<false><<<grid, 256, 0, c->stream>>>(make_entities(c), c->d_cams, c->cam0, v, g);
    return check_launch(c, "entity_primary_kernel");
}

// secondary.comp.glsl:42-48 for an entity set the kernel's compiled-in literal test does not cover
int launch_entity_shadow(uvt_ctx *c, const ViewDev &v, const GBufDev &g, uint32_t nrows, uint32_t layers) {
    if (!ent_custom(c) || !(c->params.flags & UVT_FLAG_ENTITIES)) return UVT_OK;
    const dim3 grid((c->W + 63) / 64, (nrows + 3) / 4, layers);
    entity_shadow_kernel<<<grid, 256, 0, c->stream>>>(make_entities(c), v, g);
    return check_launch(c, "entity_shadow_kernel");
}

template <int COUNT>
int launch_primary(uvt_ctx *c) {
    const ViewDev v = make_view(c, c->params.primary_max_steps);
    const GBufDev g = make_gbuf(c);
    const dim3 grid = trace_grid(c);
    if (use_pool(c)) launch_primary_pool<COUNT>(c, v, g);
    else if (COUNT != 1 && use_dense(c)) launch_primary_world<WorldDense, COUNT>(c, world_dense(c), v, g, grid);  // exact counters read the bricks
    else if (use_compact(c)) launch_primary_world<WorldCompact, COUNT>(c, world_compact(c), v, g, grid);
    else launch_primary_world<WorldRef, COUNT>(c, world_ref(c), v, g, grid);
    int rc = check_launch(c, "primary_kernel");
    if (rc != UVT_OK) return rc;
    return launch_entity_primary(c, v, g, storage_rows(c->H, c->band_rows, c->n_parts, c->part), c->layers);
}

// (re)compute the sun clearance of the block columns whose reach covers the rectangle of changed column tops
void update_sun(uvt_ctx *c, int x0, int z0, int x1, int z1) {
    const int dim = (int)c->dim;
    // top2[c] reads the tops of columns c + [0, 1]^2; sun1[X] reads top2 of X + [0, R + 1]
    const int tx0 = std::max(x0 - 1, 0), tz0 = std::max(z0 - 1, 0), tx1 = std::min(x1, dim - 1), tz1 = std::min(z1, dim - 1);
    const int nt = (tx1 - tx0 + 1) * (tz1 - tz0 + 1);
    top3_kernel<<<(nt + 255) / 256, 256, 0, c->stream>>>(c->d_tops32, c->d_top3, dim, tx0, tz0, tx1, tz1);
    const int R = sun_reach_columns((int)c->sun_steps) + 1;
    const int sx0 = std::max(tx0 - R, 0), sz0 = std::max(tz0 - R, 0);
    const int ns = (tx1 - sx0 + 1) * (tz1 - sz0 + 1);
    sun_clear_kernel<<<(ns + 255) / 256, 256, 0, c->stream>>>(c->d_top3, c->d_sun1, dim, (int)c->sun_steps, sx0, sz0, tx1, tz1);
    c->launches += 2;
}

// the sun clearance must cover the shadow step cap in force (it is built at commit time for the cap of that moment)
void ensure_sun(uvt_ctx *c) {
    if (!use_compact(c) || c->params.shadow_max_steps <= c->sun_steps) return;
    c->sun_steps = c->params.shadow_max_steps;
    update_sun(c, 0, 0, (int)c->dim - 1, (int)c->dim - 1);
}

template <int COUNT>
int launch_secondary(uvt_ctx *c) {
    ensure_sun(c);
    const ViewDev v = make_view(c, c->params.shadow_max_steps);
    const GBufDev g = make_gbuf(c);
    const dim3 grid = trace_grid(c);
    if (use_pool(c)) secondary_pool_kernel<COUNT><<<pool_grid(c), kThreads, 0, c->stream>>>(world_compact(c), v, g, c->d_counters);
    else if (COUNT != 1 && use_dense(c)) secondary_kernel<WorldDense, COUNT><<<grid, kTileThreads, 0, c->stream>>>(world_dense(c), v, g, c->d_counters);
    else if (use_compact(c)) secondary_kernel<WorldCompact, COUNT><<<grid, kTileThreads, 0, c->stream>>>(world_compact(c), v, g, c->d_counters);
    else secondary_kernel<WorldRef, COUNT><<<grid, kTileThreads, 0, c->stream>>>(world_ref(c), v, g, c->d_counters);
    int rc = check_launch(c, "secondary_kernel");
    if (rc != UVT_OK) return rc;
    return launch_entity_shadow(c, v, g, storage_rows(c->H, c->band_rows, c->n_parts, c->part), c->layers);
}

// secondary pass + blit in one launch over `nrows` local rows from v.row0 (the frame paths, when no entity pass sits between them)
bool fuse_shade(const uvt_ctx *c) { return !use_pool(c) && !(ent_custom(c) && (c->params.flags & UVT_FLAG_ENTITIES)); }

int launch_secondary_shade(uvt_ctx *c, const ViewDev &v, const GBufDev &g, dim3 grid) {
    ensure_sun(c);
    const FrameTarget ft = make_target(c);
    if (use_dense(c)) secondary_shade_kernel<WorldDense><<<grid, kTileThreads, 0, c->stream>>>(world_dense(c), v, g, ft);
    else if (use_compact(c)) secondary_shade_kernel<WorldCompact><<<grid, kTileThreads, 0, c->stream>>>(world_compact(c), v, g, ft);
    else secondary_shade_kernel<WorldRef><<<grid, kTileThreads, 0, c->stream>>>(world_ref(c), v, g, ft);
    return check_launch(c, "secondary_shade_kernel");
}

// Build the B200 layout from the committed reference layout (all on the device):
// chunk distance field -> virtual bricks -> chunks2, 8-bit material bricks, block clearances.
// scratch layout of uvt_world_commit_region (32-bit words)
constexpr size_t kScrEnts = 0, kMaxBoxChunks = 4096;                 // chunk entries of the box
constexpr size_t kScrList = kScrEnts + kMaxBoxChunks, kMaxListBricks = 8192;  // bricks to repack
constexpr size_t kScrKeys = kScrList + kMaxListBricks, kLutSize = 1024;
constexpr size_t kScrVals = kScrKeys + kLutSize;                      // kLutSize bytes
constexpr size_t kScrFlags = kScrVals + kLutSize / 4;                 // 8 counters / flags
constexpr size_t kScrChanged = kScrFlags + 8, kMaxChanged = 65536;    // chunks whose chunks2 entry changed
constexpr size_t kScratchWords = kScrChanged + kMaxChanged;

// open-addressed block word -> material id table for the repack kernels
int upload_material_lut(uvt_ctx *c, uint32_t *d_keys, uint8_t *d_vals) {
    std::vector<uint32_t> keys(kLutSize, 0);
    std::vector<uint8_t> vals(kLutSize, 0);
    const uint32_t lut_mask = (uint32_t)kLutSize - 1;
    for (size_t m = 1; m < c->mat_words.size(); ++m) {
        uint32_t h = (c->mat_words[m] * 2654435761u) & lut_mask;
        while (keys[h] != 0) h = (h + 1) & lut_mask;
        keys[h] = c->mat_words[m];
        vals[h] = (uint8_t)m;
    }
    UVT_CUDA(c, cudaMemcpyAsync(d_keys, keys.data(), kLutSize * 4, cudaMemcpyHostToDevice, c->stream));
    UVT_CUDA(c, cudaMemcpyAsync(d_vals, vals.data(), kLutSize, cudaMemcpyHostToDevice, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));  // the staging vectors die at scope exit
    return UVT_OK;
}

// clear4 from the column tops, then y_clear = its maximum; (x0, z0)-(x1, z1) = the block columns whose tops changed
int finish_tops(uvt_ctx *c, unsigned int *d_max, int x0, int z0, int x1, int z1) {
    const int nq = (int)((c->dim / 4) * (c->dim / 4));
    quad_clear_kernel<<<(nq + 255) / 256, 256, 0, c->stream>>>(c->d_tops32, c->d_clear4, (int)c->dim);
    const int d4 = (int)(c->dim / 4), d16 = (d4 + 3) / 4, d64 = (d16 + 3) / 4;
    coarse_clear_kernel<<<(d16 * d16 + 127) / 128, 128, 0, c->stream>>>(c->d_clear4, c->d_clear16, d4);
    coarse_clear_kernel<<<(d64 * d64 + 127) / 128, 128, 0, c->stream>>>(c->d_clear16, c->d_clear64, d16);
    c->sun_steps = std::max<uint32_t>(std::max<uint32_t>(c->params.shadow_max_steps, c->sun_steps), 1u);
    update_sun(c, x0, z0, x1, z1);
    c->launches += 3;
    UVT_CUDA(c, cudaMemsetAsync(d_max, 0, 4, c->stream));
    max_clear_kernel<<<(nq + 255) / 256, 256, 0, c->stream>>>(c->d_clear4, nq, d_max);
    c->launches += 2;
    unsigned int top = 0;
    UVT_CUDA(c, cudaMemcpyAsync(&top, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->y_clear = (int32_t)top;
    return UVT_OK;
}

// Build the B200 layout from the committed reference layout (all on the device):
// chunk distance field -> virtual bricks -> chunks2, 8-bit material bricks, block clearances.
int build_compact(uvt_ctx *c, size_t n_bricks, size_t n_words) {
    const int cd = (int)c->cd;
    const size_t n_chunks = (size_t)cd * cd * cd;
    const size_t n2 = (size_t)(cd + 1) * (cd + 1) * (cd + 1);
    c->incremental_ok = false;
    auto cleanup = [&]() {};
#define UVT_CUDA_C(expr)                                                                                      \
    do {                                                                                                      \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            cleanup();                                                                                        \
            return set_error(c, e_ == cudaErrorMemoryAllocation ? UVT_ERR_OOM : UVT_ERR_CUDA, "%s: %s (%s:%d)", \
                             #expr, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
        }                                                                                                     \
    } while (0)
    if (!c->d_scratch) UVT_CUDA_C(cudaMalloc(&c->d_scratch, kScratchWords * 4));
    unsigned int *d_counter = reinterpret_cast<unsigned int *>(c->d_scratch + kScrFlags);
    if (!c->d_field) {
        UVT_CUDA_C(cudaMalloc(&c->d_field, n_chunks));
        UVT_CUDA_C(cudaMalloc(&c->d_field_tmp[0], n_chunks));
        UVT_CUDA_C(cudaMalloc(&c->d_field_tmp[1], n_chunks));
    }
    uint8_t *f0 = c->d_field;
    const unsigned cblocks = (unsigned)((n_chunks + 255) / 256);
    const ChunkBox whole = {0, 0, 0, cd, cd, cd};
    field_pass_x_kernel<<<cblocks, 256, 0, c->stream>>>(c->d_chunks, c->d_field_tmp[0], cd, whole);
    field_pass_kernel<<<cblocks, 256, 0, c->stream>>>(c->d_field_tmp[0], c->d_field_tmp[1], cd, 1, whole);
    field_pass_kernel<<<cblocks, 256, 0, c->stream>>>(c->d_field_tmp[1], f0, cd, 2, whole);
    UVT_CUDA_C(cudaMemsetAsync(d_counter, 0, 4, c->stream));
    count_virtual_kernel<<<cblocks, 256, 0, c->stream>>>(c->d_chunks, f0, n_chunks, d_counter);
    unsigned int n_virtual = 0;
    UVT_CUDA_C(cudaMemcpyAsync(&n_virtual, d_counter, 4, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA_C(cudaStreamSynchronize(c->stream));
    c->launches += 4;
    // spare real-brick slots below the virtual bricks, so that later edits can add bricks without renumbering
    const size_t v_base = n_bricks + std::max<size_t>(256, n_bricks / 16);
    const size_t n_total = v_base + n_virtual;
    if (n_total * 512 >= (1ull << 32)) {  // brick byte offsets are 32-bit in the traversal kernel
        cleanup();
        c->compact_ok = false;
        return UVT_OK;
    }
    if (n_total > c->d_brick8_capacity) {
        cudaFree(c->d_bricks8); cudaFree(c->d_rowmask); cudaFree(c->d_brick_chunk);
        c->d_bricks8 = nullptr; c->d_rowmask = nullptr; c->d_brick_chunk = nullptr;
        c->d_brick8_capacity = 0;
        const size_t want = std::min<size_t>(n_total + n_virtual / 4 + 256, ((1ull << 32) - 1) / 512);
        UVT_CUDA_C(cudaMalloc(&c->d_bricks8, want * 512));
        UVT_CUDA_C(cudaMalloc(&c->d_rowmask, want * 64));
        UVT_CUDA_C(cudaMalloc(&c->d_brick_chunk, want * 4));
        c->d_brick8_capacity = want;
    }
    if (!c->d_tops32) UVT_CUDA_C(cudaMalloc(&c->d_tops32, (size_t)c->dim * c->dim * 4));
    UVT_CUDA_C(cudaMemsetAsync(c->d_brick_chunk, 0xFF, c->d_brick8_capacity * 4, c->stream));  // kNoChunk
    UVT_CUDA_C(cudaMemsetAsync(d_counter, 0, 4, c->stream));
    build_chunks2_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, c->stream>>>(c->d_chunks, f0, c->d_chunks2, c->d_brick_chunk, cd, (uint32_t)v_base, d_counter);
    c->launches++;
    UVT_CUDA_C(cudaMemsetAsync(c->d_bricks8, 0, n_total * 512, c->stream));
    if (n_bricks) {
        uint32_t *d_keys = c->d_scratch + kScrKeys;
        uint8_t *d_vals = reinterpret_cast<uint8_t *>(c->d_scratch + kScrVals);
        int rc = upload_material_lut(c, d_keys, d_vals);
        if (rc != UVT_OK) { cleanup(); return rc; }
        repack_bricks_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_bricks, c->d_bricks8, n_words, d_keys, d_vals, (uint32_t)kLutSize - 1);
        c->launches++;
    }
    {   // column tops -> dilated 4x4 column-group tops (sky_sealed) -> y_clear
        UVT_CUDA_C(cudaMemsetAsync(c->d_tops32, 0, (size_t)c->dim * c->dim * 4, c->stream));
        if (n_bricks) {
            column_tops_kernel<<<(unsigned)n_bricks, 64, 0, c->stream>>>(c->d_bricks8, c->d_brick_chunk, cd, c->d_tops32);
            c->launches++;
        }
        int rc = finish_tops(c, d_counter, 0, 0, (int)c->dim - 1, (int)c->dim - 1);
        if (rc != UVT_OK) { cleanup(); return rc; }
    }
    {
        const size_t n_rows = n_total * 64;
        brick_rowmask_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, c->stream>>>(c->d_bricks8, n_rows, c->d_rowmask);
        clearance_kernel<<<(unsigned)n_total, 512, 0, c->stream>>>(c->d_chunks2, cd, c->d_brick_chunk, c->d_rowmask, c->d_bricks8);
        c->launches += 2;
    }
    // dense block grid (skipped when dim^3 bytes do not fit comfortably: the brick path is used instead)
    c->dense_valid = false;
    {
        const size_t bytes = (size_t)c->dim * c->dim * c->dim;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (!c->d_dense && bytes <= (size_t)16 << 30 && bytes * 2 < free_b) {
            if (cudaMalloc(&c->d_dense, bytes) != cudaSuccess) { c->d_dense = nullptr; (void)cudaGetLastError(); }
        }
        if (c->d_dense) {
            dense_fill_kernel<<<(unsigned)n_chunks, 64, 0, c->stream>>>(c->d_chunks2, c->d_bricks8, c->d_dense, cd);
            c->launches++;
            c->dense_valid = true;
        }
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
#undef UVT_CUDA_C
    if (e != cudaSuccess) return set_error(c, UVT_ERR_CUDA, "building the compact layout failed: %s", cudaGetErrorString(e));
    c->n_total_bricks = n_total;
    c->v_base = v_base;
    c->n_virtual = n_virtual;
    c->incremental_ok = true;
    return UVT_OK;
}

}  // namespace

extern "C" {

void uvt_default_params(uvt_params *p) {
    std::memset(p, 0, sizeof *p);
    p->map_dim = 512;
    p->primary_max_steps = 192;
    p->shadow_max_steps = 48;
    p->edit_max_steps = 64;
    p->epsilon = 0.001f;
    p->flags = UVT_FLAG_ENTITIES;
    p->layout = UVT_LAYOUT_COMPACT;
    p->scheduler = UVT_SCHED_TILE;
}

int uvt_abi_version(void) { return UVT_ABI_VERSION; }

int uvt_create(const uvt_params *params, int device, uvt_ctx **out) {
    if (!out) return set_error(nullptr, UVT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(nullptr, UVT_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU path",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return set_error(nullptr, UVT_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(nullptr, UVT_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_error(nullptr, UVT_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return set_error(nullptr, UVT_ERR_NO_DEVICE, "device %d is sm_%d%d; this build carries sm_100a code only", device, prop.major, prop.minor);

    uvt_params prm;
    if (params) prm = *params;
    else uvt_default_params(&prm);
    // the same checks the setters make (uvt_set_max_steps, uvt_set_layout, uvt_set_scheduler)
    if (prm.primary_max_steps > 65535u || prm.shadow_max_steps > 65535u || prm.edit_max_steps > 65535u)
        return set_error(nullptr, UVT_ERR_INVALID, "step caps must fit 16 bits (uvt_hit.trips)");
    if (prm.layout != UVT_LAYOUT_COMPACT && prm.layout != UVT_LAYOUT_REFERENCE) return set_error(nullptr, UVT_ERR_INVALID, "unknown layout %u", prm.layout);
    if (prm.scheduler != UVT_SCHED_POOL && prm.scheduler != UVT_SCHED_TILE) return set_error(nullptr, UVT_ERR_INVALID, "unknown scheduler %u", prm.scheduler);
    if (prm.map_dim != 0 && (prm.map_dim % 8u != 0 || prm.map_dim > 4096u)) return set_error(nullptr, UVT_ERR_INVALID, "map_dim must be a multiple of 8, at most 4096");
    if (!(prm.epsilon == prm.epsilon)) return set_error(nullptr, UVT_ERR_INVALID, "epsilon is NaN");

    uvt_ctx *c = new uvt_ctx;
    c->params = prm;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    auto fail = [&](cudaError_t err, const char *what) {
        set_error(nullptr, UVT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        uvt_destroy(c);
        return UVT_ERR_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    c->stream = c->own_stream;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j)
            if ((e = cudaEventCreate(&c->ev[i][j])) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate (copy)");
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaEventCreateWithFlags(&c->snap_ready[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
        if ((e = cudaEventCreateWithFlags(&c->snap_done[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
    }
    if ((e = cudaMalloc(&c->d_counters, sizeof(DevCounters))) != cudaSuccess) return fail(e, "cudaMalloc counters");
    if ((e = cudaMalloc(&c->d_pick, 32)) != cudaSuccess) return fail(e, "cudaMalloc pick");
    if ((e = cudaMalloc(&c->d_sink, 4)) != cudaSuccess) return fail(e, "cudaMalloc sink");
    if ((e = cudaMalloc(&c->d_mat_word, 256 * 4)) != cudaSuccess) return fail(e, "cudaMalloc mat_word");
    if ((e = cudaMalloc(&c->d_mat_color, 256 * 512 * 4)) != cudaSuccess) return fail(e, "cudaMalloc mat_color");
    if ((e = cudaMalloc(&c->d_mat_mask, 256 * 16 * 4)) != cudaSuccess) return fail(e, "cudaMalloc mat_mask");
    std::memset(&c->cam0, 0, sizeof c->cam0);
    *out = c;
    return UVT_OK;
}

void uvt_destroy(uvt_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    uvt_nccl_shutdown(c);
    cudaFree(c->pgen.vh); cudaFree(c->pgen.seeds); cudaFree(c->pgen.deco); cudaFree(c->pgen.trees);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->snap[i]);
        if (c->snap_ready[i]) cudaEventDestroy(c->snap_ready[i]);
        if (c->snap_done[i]) cudaEventDestroy(c->snap_done[i]);
    }
    free_gbuffer(c);
    for (void *p : c->pinned) cudaFreeHost(p);
    if (!c->staging_borrowed) { cudaFreeHost(c->h_chunks); cudaFreeHost(c->h_bricks); }
    cudaFree(c->d_chunks); cudaFree(c->d_bricks); cudaFree(c->d_bricks8); cudaFree(c->d_models); cudaFree(c->d_chunks2); cudaFree(c->d_clear4); cudaFree(c->d_dense);
    cudaFree(c->d_rowmask); cudaFree(c->d_brick_chunk); cudaFree(c->d_tops32); cudaFree(c->d_scratch);
    cudaFree(c->d_field); cudaFree(c->d_field_tmp[0]); cudaFree(c->d_field_tmp[1]); cudaFree(c->d_clear64); cudaFree(c->d_clear16); cudaFree(c->d_sun1); cudaFree(c->d_top3);
    cudaFree(c->d_mat_word); cudaFree(c->d_mat_color); cudaFree(c->d_mat_mask);
    cudaFree(c->d_cams); cudaFree(c->d_counters); cudaFree(c->d_pick); cudaFree(c->d_sink); cudaFree(c->shared_frame); cudaFree(c->d_ent_model);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j)
            if (c->ev[i][j]) cudaEventDestroy(c->ev[i][j]);
    if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    if (c->join_ev) cudaEventDestroy(c->join_ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char *uvt_last_error(uvt_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int uvt_set_stream(uvt_ctx *c, void *cuda_stream) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return UVT_OK;
}

int uvt_get_params(uvt_ctx *c, uvt_params *out) {
    if (!c || !out) return UVT_ERR_INVALID;
    *out = c->params;
    return UVT_OK;
}

int uvt_set_layout(uvt_ctx *c, uint32_t layout) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, layout == UVT_LAYOUT_COMPACT || layout == UVT_LAYOUT_REFERENCE, "unknown layout");
    c->params.layout = layout;
    return UVT_OK;
}

int uvt_set_scheduler(uvt_ctx *c, uint32_t scheduler) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, scheduler == UVT_SCHED_POOL || scheduler == UVT_SCHED_TILE, "unknown scheduler");
    c->params.scheduler = scheduler;
    return UVT_OK;
}

int uvt_effective_layout(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    return use_compact(c) ? (int)UVT_LAYOUT_COMPACT : (int)UVT_LAYOUT_REFERENCE;
}

int uvt_set_max_steps(uvt_ctx *c, uint32_t primary, uint32_t shadow) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, primary <= 65535u && shadow <= 65535u, "step caps must fit 16 bits (uvt_hit.trips)");
    c->params.primary_max_steps = primary;
    c->params.shadow_max_steps = shadow;
    return UVT_OK;
}

// ---- pipelines -----------------------------------------------------------------------------
int uvt_pipeline_create(uvt_ctx *c, uvt_pipeline_kind kind, uvt_pipeline **out) {
    if (!c || !out) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, kind >= UVT_PIPELINE_PRIMARY && kind <= UVT_PIPELINE_BLIT, "unknown pipeline kind");
    // the analogue of compile + link: make sure the kernel image for this device resolves
    cudaFuncAttributes fa;
    const void *fn = nullptr;
    switch (kind) {
        case UVT_PIPELINE_PRIMARY: fn = (const void *)primary_kernel<WorldCompact, 0, false, false>; break;
        case UVT_PIPELINE_SECONDARY: fn = (const void *)secondary_kernel<WorldCompact, 0>; break;
        case UVT_PIPELINE_EDIT: fn = (const void *)pick_kernel<WorldCompact>; break;
        case UVT_PIPELINE_BLIT: fn = (const void *)shade_kernel; break;
    }
    UVT_CUDA(c, cudaFuncGetAttributes(&fa, fn));
    *out = new uvt_pipeline{c, kind};
    return UVT_OK;
}

void uvt_pipeline_destroy(uvt_pipeline *p) { delete p; }

int uvt_pipeline_dispatch(uvt_pipeline *p, uint32_t gx, uint32_t gy, uint32_t gz) {
    if (!p) return UVT_ERR_INVALID;
    uvt_ctx *c = p->ctx;
    if (p->kind == UVT_PIPELINE_BLIT) return uvt_shade(c);  // RasterPipeline.draw(4)
    if (p->kind == UVT_PIPELINE_EDIT) {
        UVT_REQUIRE(c, gx == 1 && gy == 1 && gz == 1, "terrain_edit dispatches 1x1x1");
        uvt_hit h;
        return uvt_pick(c, &h);
    }
    // game.zig:241-242: (W/32 + 1) x (H/32 + 1) x 1 groups of 32x32 threads
    UVT_REQUIRE(c, gz == 1 && (uint64_t)gx * 32u >= c->W && (uint64_t)gy * 32u >= c->H,
                "dispatch does not cover the G-buffer (expected (W/32+1) x (H/32+1) x 1 groups)");
    return p->kind == UVT_PIPELINE_PRIMARY ? uvt_dispatch_primary(c) : uvt_dispatch_secondary(c);
}

// ---- world ------------------------------------------------------------------------------------
// drop the current world and size the device-side tables for `dim` (host staging is set by the caller)
static int reset_world(uvt_ctx *c, uint32_t dim) {
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!c->staging_borrowed) { cudaFreeHost(c->h_chunks); cudaFreeHost(c->h_bricks); }
    cudaFree(c->d_chunks); cudaFree(c->d_chunks2); cudaFree(c->d_clear4); cudaFree(c->d_dense); cudaFree(c->d_tops32); cudaFree(c->d_clear64); cudaFree(c->d_clear16); cudaFree(c->d_sun1); cudaFree(c->d_top3);
    c->d_clear64 = c->d_clear16 = c->d_sun1 = c->d_top3 = nullptr;
    cudaFree(c->d_field); cudaFree(c->d_field_tmp[0]); cudaFree(c->d_field_tmp[1]);
    c->d_field = c->d_field_tmp[0] = c->d_field_tmp[1] = nullptr;
    c->d_dense = nullptr;
    c->d_tops32 = nullptr;
    c->incremental_ok = false;
    c->dense_valid = false;
    c->h_chunks = c->h_bricks = nullptr;
    c->staging_borrowed = false;
    c->d_chunks = nullptr;
    c->d_chunks2 = nullptr;
    c->d_clear4 = nullptr;
    c->dim = dim;
    c->cd = dim / 8;
    c->params.map_dim = dim;
    c->h_capacity = 0;
    c->n_bricks = 0;
    c->world_committed = false;
    const size_t n_chunks = (size_t)c->cd * c->cd * c->cd;
    UVT_CUDA(c, cudaMalloc(&c->d_chunks, n_chunks * 4));
    UVT_CUDA(c, cudaMalloc(&c->d_chunks2, (size_t)(c->cd + 1) * (c->cd + 1) * (c->cd + 1) * 4));
    UVT_CUDA(c, cudaMalloc(&c->d_clear4, (size_t)(dim / 4) * (dim / 4) * 2));
    UVT_CUDA(c, cudaMalloc(&c->d_clear16, (size_t)((dim + 15) / 16) * ((dim + 15) / 16) * 2));
    UVT_CUDA(c, cudaMalloc(&c->d_sun1, (size_t)dim * dim * 2));
    UVT_CUDA(c, cudaMalloc(&c->d_top3, (size_t)dim * dim * 2));
    UVT_CUDA(c, cudaMalloc(&c->d_clear64, (size_t)((dim + 63) / 64) * ((dim + 63) / 64) * 2));
    return UVT_OK;
}

int uvt_world_alloc(uvt_ctx *c, uint32_t dim, uint32_t **chunks_host, uint32_t **bricks_host, size_t brick_capacity) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, chunks_host && bricks_host, "NULL out pointer");
    UVT_REQUIRE(c, dim >= 8 && dim % 8 == 0 && dim <= 4096, "dim must be a multiple of 8 in [8, 4096]");
    UVT_REQUIRE(c, brick_capacity > 0, "brick_capacity must be > 0");
    int rc = reset_world(c, dim);
    if (rc != UVT_OK) return rc;
    const size_t n_chunks = (size_t)c->cd * c->cd * c->cd;
    UVT_CUDA(c, cudaHostAlloc(&c->h_chunks, n_chunks * 4, cudaHostAllocPortable));
    UVT_CUDA(c, cudaHostAlloc(&c->h_bricks, brick_capacity * 2048, cudaHostAllocPortable));
    std::memset(c->h_chunks, 0, n_chunks * 4);              // voxel.zig:34
    std::memset(c->h_bricks, 0, brick_capacity * 2048);     // GL zero-initialised storage
    c->h_capacity = brick_capacity;
    *chunks_host = c->h_chunks;
    *bricks_host = c->h_bricks;
    return UVT_OK;
}

int uvt_world_use_staging(uvt_ctx *c, uint32_t dim, uint32_t *chunks_host, uint32_t *bricks_host, size_t brick_capacity) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, bricks_host && brick_capacity > 0, "NULL brick pool");
    if (dim == 0) {  // the caller's pool moved or grew: same world, new pool address
        UVT_REQUIRE(c, c->staging_borrowed && c->h_chunks, "no caller-owned staging to update");
        UVT_REQUIRE(c, brick_capacity >= c->h_capacity, "cannot shrink the brick pool (buffer.zig:51-52)");
        UVT_CUDA(c, cudaStreamSynchronize(c->stream));
        c->h_bricks = bricks_host;
        c->h_capacity = brick_capacity;
        return UVT_OK;
    }
    UVT_REQUIRE(c, chunks_host, "NULL chunk table");
    UVT_REQUIRE(c, dim >= 8 && dim % 8 == 0 && dim <= 4096, "dim must be a multiple of 8 in [8, 4096]");
    int rc = reset_world(c, dim);
    if (rc != UVT_OK) return rc;
    c->h_chunks = chunks_host;
    c->h_bricks = bricks_host;
    c->h_capacity = brick_capacity;
    c->staging_borrowed = true;
    return UVT_OK;
}

int uvt_world_grow(uvt_ctx *c, size_t new_capacity, uint32_t **bricks_host) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->h_bricks && bricks_host, "no world allocated");
    UVT_REQUIRE(c, !c->staging_borrowed, "the staging belongs to the caller (uvt_world_use_staging): grow it there");
    UVT_REQUIRE(c, new_capacity >= c->h_capacity, "cannot shrink the brick pool (buffer.zig:51-52)");
    if (new_capacity == c->h_capacity) { *bricks_host = c->h_bricks; return UVT_OK; }
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));  // an upload may still be reading the old staging
    uint32_t *nb = nullptr;
    UVT_CUDA(c, cudaHostAlloc(&nb, new_capacity * 2048, cudaHostAllocPortable));
    std::memcpy(nb, c->h_bricks, c->h_capacity * 2048);  // copyNamedBufferSubData (buffer.zig:57)
    std::memset((uint8_t *)nb + c->h_capacity * 2048, 0, (new_capacity - c->h_capacity) * 2048);
    cudaFreeHost(c->h_bricks);
    c->h_bricks = nb;
    c->h_capacity = new_capacity;
    *bricks_host = nb;
    return UVT_OK;
}

int uvt_world_commit(uvt_ctx *c, size_t n_bricks) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->h_chunks && c->h_bricks, "no world allocated");
    UVT_REQUIRE(c, n_bricks <= c->h_capacity, "n_bricks exceeds the pool capacity");
    const size_t n_chunks = (size_t)c->cd * c->cd * c->cd;
    const bool trace = std::getenv("UVT_TRACE_COMMIT") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    auto t_phase = now();
    // every chunk entry must name a committed brick
    // (checked on the host: a bad index would be an out-of-bounds device read)
    for (size_t i = 0; i < n_chunks; ++i)
        if (c->h_chunks[i] > n_bricks) return set_error(c, UVT_ERR_INVALID, "chunk entry %zu names brick %u >= n_bricks %zu", i, c->h_chunks[i] - 1, n_bricks);

    c->world_committed = false;  // until every step below has succeeded
    const size_t cap = std::max<size_t>(n_bricks, 1);
    if (cap > c->d_brick_capacity) {
        cudaFree(c->d_bricks);
        c->d_bricks = nullptr;
        const size_t want = std::max(cap, c->h_capacity);
        UVT_CUDA(c, cudaMalloc(&c->d_bricks, want * 2048));
        c->d_brick_capacity = want;
    }
    UVT_CUDA(c, cudaMemcpyAsync(c->d_chunks, c->h_chunks, n_chunks * 4, cudaMemcpyHostToDevice, c->stream));
    if (n_bricks) UVT_CUDA(c, cudaMemcpyAsync(c->d_bricks, c->h_bricks, n_bricks * 2048, cudaMemcpyHostToDevice, c->stream));

    if (trace) { cudaStreamSynchronize(c->stream); std::fprintf(stderr, "[uvt commit] validate + upload %.1f ms\n", ms_since(t_phase)); t_phase = now(); }
    // material table: the distinct block words, ascending (collected on the device from the uploaded bricks)
    c->mat_words.assign(1, 0u);
    const size_t n_words = n_bricks * 512;
    bool overflow = false;
    if (n_words) {
        uint32_t *d_set = nullptr;
        UVT_CUDA(c, cudaMalloc(&d_set, (kWordSetSlots + 1) * 4));
        cudaError_t e = cudaMemsetAsync(d_set, 0, (kWordSetSlots + 1) * 4, c->stream);
        distinct_words_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->d_bricks, n_words, d_set, d_set + kWordSetSlots);
        c->launches++;
        std::vector<uint32_t> set(kWordSetSlots + 1);
        if (e == cudaSuccess) e = cudaMemcpyAsync(set.data(), d_set, set.size() * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        cudaFree(d_set);
        if (e != cudaSuccess) return set_error(c, UVT_ERR_CUDA, "collecting the block words failed: %s", cudaGetErrorString(e));
        overflow = set[kWordSetSlots] != 0;
        for (uint32_t k = 0; k < kWordSetSlots && !overflow; ++k) {
            if (set[k] == 0) continue;
            if (c->mat_words.size() >= kMatLimit) { overflow = true; break; }  // ids >= kMatLimit encode clearances
            c->mat_words.push_back(set[k]);
        }
        if (overflow) c->mat_words.resize(1);
    }
    std::sort(c->mat_words.begin() + 1, c->mat_words.end());  // canonical ids: the same materials always get the same ids
    c->compact_ok = !overflow;
    c->incremental_ok = false;
    if (trace) { std::fprintf(stderr, "[uvt commit] material scan %.1f ms\n", ms_since(t_phase)); t_phase = now(); }
    if (c->compact_ok) {
        int rc = build_compact(c, n_bricks, n_words);
        if (rc != UVT_OK) return rc;
    }
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (trace) std::fprintf(stderr, "[uvt commit] build_compact %.1f ms\n", ms_since(t_phase));
    c->n_bricks = n_bricks;
    c->world_committed = true;
    c->materials_dirty = true;
    return UVT_OK;
}

// Incremental publish of a block box.  Everything a full commit derives from the edited
// blocks is refreshed in place: the bricks of the box, the clearances and dense-grid bytes of every chunk
// within two chunks of it (clearances look kClearCap = 16 blocks far), the column tops under it; a new brick
// additionally refreshes the chunk distance field and gives newly adjacent empty chunks a virtual brick.
static int commit_region_impl(uvt_ctx *c, size_t n_bricks, const uint32_t lo[3], const uint32_t hi[3]);

int uvt_world_commit_region(uvt_ctx *c, size_t n_bricks, const uint32_t lo[3], const uint32_t hi[3]) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    const int rc = commit_region_impl(c, n_bricks, lo, hi);
    if (rc != UVT_OK && rc != UVT_ERR_INVALID) {
        // a device-side failure half way leaves the derived layout inconsistent: nothing may be traced until a full commit succeeds
        c->incremental_ok = false;
        c->world_committed = false;
    }
    return rc;
}

static int commit_region_impl(uvt_ctx *c, size_t n_bricks, const uint32_t lo[3], const uint32_t hi[3]) {
    UVT_REQUIRE(c, c->h_chunks && c->h_bricks, "no world allocated");
    UVT_REQUIRE(c, lo && hi, "NULL box");
    UVT_REQUIRE(c, n_bricks <= c->h_capacity, "n_bricks exceeds the pool capacity");
    for (int a = 0; a < 3; ++a) UVT_REQUIRE(c, lo[a] <= hi[a] && hi[a] < c->dim, "box must satisfy lo <= hi < dim");
    const int cd = (int)c->cd;
    const int o[3] = {(int)(lo[0] >> 3), (int)(lo[1] >> 3), (int)(lo[2] >> 3)};
    const int bx = (int)(hi[0] >> 3) - o[0] + 1, by = (int)(hi[1] >> 3) - o[1] + 1, bz = (int)(hi[2] >> 3) - o[2] + 1;
    const size_t nb = (size_t)bx * by * bz;
    const size_t old_n = c->n_bricks;
    // anything the in-place path cannot express is published by a full commit
    if (!c->world_committed || !c->compact_ok || !c->incremental_ok || n_bricks < old_n || n_bricks > c->v_base ||
        n_bricks > c->d_brick_capacity || nb > kMaxBoxChunks || n_bricks - old_n > kMaxListBricks / 2)
        return uvt_world_commit(c, n_bricks);

    // ---- host: box entries, bricks to refresh, new materials
    std::vector<uint32_t> ents(nb), list;
    for (int z = 0; z < bz; ++z)
        for (int y = 0; y < by; ++y)
            for (int x = 0; x < bx; ++x) {
                const size_t j = (size_t)(o[0] + x) + (size_t)cd * ((size_t)(o[1] + y) + (size_t)(o[2] + z) * cd);
                const uint32_t e = c->h_chunks[j];
                if (e > n_bricks) return set_error(c, UVT_ERR_INVALID, "chunk entry %zu names brick %u >= n_bricks %zu", j, e - 1, n_bricks);
                ents[(size_t)x + (size_t)bx * ((size_t)y + (size_t)by * z)] = e;
                if (e != 0 && e - 1 < old_n) list.push_back(e - 1);
            }
    for (size_t b = old_n; b < n_bricks; ++b) list.push_back((uint32_t)b);
    std::sort(list.begin(), list.end());
    list.erase(std::unique(list.begin(), list.end()), list.end());
    if (list.size() > kMaxListBricks) return uvt_world_commit(c, n_bricks);
    bool new_materials = false;
    for (uint32_t b : list) {
        const uint32_t *bw = c->h_bricks + (size_t)b * 512;
        uint32_t last_word = 0;
        for (int i = 0; i < 512; ++i) {
            const uint32_t wd = bw[i];
            if (wd == 0 || wd == last_word) continue;
            last_word = wd;
            if (std::find(c->mat_words.begin() + 1, c->mat_words.end(), wd) == c->mat_words.end()) {
                if (c->mat_words.size() >= kMatLimit) return uvt_world_commit(c, n_bricks);  // the full commit falls back to the reference layout
                c->mat_words.push_back(wd);
                new_materials = true;
            }
        }
    }

    // ---- uploads
    uint32_t *d_ents = c->d_scratch + kScrEnts, *d_list = c->d_scratch + kScrList, *d_keys = c->d_scratch + kScrKeys;
    uint8_t *d_vals = reinterpret_cast<uint8_t *>(c->d_scratch + kScrVals);
    unsigned int *d_flags = reinterpret_cast<unsigned int *>(c->d_scratch + kScrFlags);
    uint32_t *d_changed = c->d_scratch + kScrChanged;
    for (uint32_t b : list)
        UVT_CUDA(c, cudaMemcpyAsync(c->d_bricks + (size_t)b * 512, c->h_bricks + (size_t)b * 512, 2048, cudaMemcpyHostToDevice, c->stream));
    UVT_CUDA(c, cudaMemcpyAsync(d_ents, ents.data(), nb * 4, cudaMemcpyHostToDevice, c->stream));
    if (!list.empty()) UVT_CUDA(c, cudaMemcpyAsync(d_list, list.data(), list.size() * 4, cudaMemcpyHostToDevice, c->stream));
    UVT_CUDA(c, cudaMemsetAsync(d_flags, 0, 32, c->stream));
    int rc = upload_material_lut(c, d_keys, d_vals);
    if (rc != UVT_OK) return rc;

    // ---- chunk entries and bricks
    apply_chunk_box_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, c->stream>>>(c->d_chunks, d_ents, cd, o[0], o[1], o[2], bx, by, bz, d_flags + 0);
    c->launches++;
    if (!list.empty()) {
        repack_list_kernel<<<(unsigned)list.size(), 64, 0, c->stream>>>(c->d_bricks, c->d_bricks8, c->d_rowmask, d_list, d_keys, d_vals, (uint32_t)kLutSize - 1);
        c->launches++;
    }
    unsigned int table_changed = 0;
    UVT_CUDA(c, cudaMemcpyAsync(&table_changed, d_flags + 0, 4, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_bricks = n_bricks;
    if (new_materials) c->materials_dirty = true;

    unsigned int n_changed = 0;
    if (table_changed) {
        // distance field -> new virtual bricks -> chunks2.  A distance depends on the chunks within kFieldCap - 1 of it on
        // every axis, so only the box grown by R = kFieldCap can change; the separable passes need their inputs one R further
        // along the axes still to come.
        const int R = kFieldCap;
        const int ext[3] = {bx, by, bz};
        auto grown = [&](int rx, int ry, int rz) {
            const int r[3] = {rx, ry, rz};
            int lo3[3], hi3[3];
            for (int a = 0; a < 3; ++a) {
                lo3[a] = std::max(o[a] - r[a], 0);
                hi3[a] = std::min(o[a] + ext[a] - 1 + r[a], cd - 1);
            }
            return ChunkBox{lo3[0], lo3[1], lo3[2], hi3[0] - lo3[0] + 1, hi3[1] - lo3[1] + 1, hi3[2] - lo3[2] + 1};
        };
        auto blocks_of = [](const ChunkBox &b) { return (unsigned)((b.count() + 255) / 256); };
        const ChunkBox bx_ = grown(R, 2 * R, 2 * R), by_ = grown(R, R, 2 * R), bz_ = grown(R, R, R);
        field_pass_x_kernel<<<blocks_of(bx_), 256, 0, c->stream>>>(c->d_chunks, c->d_field_tmp[0], cd, bx_);
        field_pass_kernel<<<blocks_of(by_), 256, 0, c->stream>>>(c->d_field_tmp[0], c->d_field_tmp[1], cd, 1, by_);
        field_pass_kernel<<<blocks_of(bz_), 256, 0, c->stream>>>(c->d_field_tmp[1], c->d_field, cd, 2, bz_);
        count_new_virtual_kernel<<<blocks_of(bz_), 256, 0, c->stream>>>(c->d_chunks, c->d_field, c->d_chunks2, cd, (uint32_t)c->v_base, d_flags + 1, bz_);
        c->launches += 4;
        unsigned int n_new = 0;
        cudaError_t e = cudaMemcpyAsync(&n_new, d_flags + 1, 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return set_error(c, UVT_ERR_CUDA, "distance field: %s", cudaGetErrorString(e));
        if (c->v_base + c->n_virtual + n_new > c->d_brick8_capacity) return uvt_world_commit(c, n_bricks);  // out of virtual-brick slots: rebuild
        unsigned int nv = (unsigned int)c->n_virtual;
        cudaMemcpyAsync(d_flags + 2, &nv, 4, cudaMemcpyHostToDevice, c->stream);
        update_chunks2_kernel<<<blocks_of(bz_), 256, 0, c->stream>>>(c->d_chunks, c->d_field, c->d_chunks2, c->d_brick_chunk, c->d_bricks8, c->d_rowmask, cd,
                                                                    (uint32_t)c->v_base, d_flags + 2, d_changed, (uint32_t)kMaxChanged, d_flags + 3, bz_);
        c->launches++;
        unsigned int out[2] = {0, 0};
        e = cudaMemcpyAsync(out, d_flags + 2, 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return set_error(c, UVT_ERR_CUDA, "chunks2 update: %s", cudaGetErrorString(e));
        c->n_virtual = out[0];
        c->n_total_bricks = c->v_base + c->n_virtual;
        n_changed = out[1];
    }

    // ---- column tops under the box -> clear4 -> y_clear
    UVT_CUDA(c, cudaMemset2DAsync(c->d_tops32 + (size_t)o[0] * 8 + (size_t)c->dim * ((size_t)o[2] * 8), (size_t)c->dim * 4, 0, (size_t)bx * 8 * 4,
                                  (size_t)bz * 8, c->stream));
    column_tops_box_kernel<<<dim3((unsigned)bx, (unsigned)cd, (unsigned)bz), 64, 0, c->stream>>>(c->d_chunks, c->d_bricks8, cd, o[0], o[2], c->d_tops32);
    c->launches++;
    rc = finish_tops(c, d_flags + 4, o[0] * 8, o[2] * 8, (o[0] + bx) * 8 - 1, (o[2] + bz) * 8 - 1);
    if (rc != UVT_OK) return rc;

    // ---- clearances and dense bytes of the box grown by two chunks
    int g0[3], g1[3];
    const int ext[3] = {bx, by, bz};
    for (int a = 0; a < 3; ++a) {
        g0[a] = std::max(o[a] - 2, 0);
        g1[a] = std::min(o[a] + ext[a] - 1 + 2, cd - 1);
    }
    const dim3 ggrid((unsigned)(g1[0] - g0[0] + 1), (unsigned)(g1[1] - g0[1] + 1), (unsigned)(g1[2] - g0[2] + 1));
    clearance_box_kernel<<<ggrid, 512, 0, c->stream>>>(c->d_chunks2, cd, g0[0], g0[1], g0[2], c->d_rowmask, c->d_bricks8);
    c->launches++;
    if (c->d_dense && c->dense_valid) {
        if (n_changed > kMaxChanged) {
            dense_fill_kernel<<<(unsigned)((size_t)cd * cd * cd), 64, 0, c->stream>>>(c->d_chunks2, c->d_bricks8, c->d_dense, cd);
        } else {
            dense_fill_box_kernel<<<ggrid, 64, 0, c->stream>>>(c->d_chunks2, c->d_bricks8, c->d_dense, cd, g0[0], g0[1], g0[2]);
            if (n_changed) {
                dense_fill_list_kernel<<<n_changed, 64, 0, c->stream>>>(c->d_chunks2, c->d_bricks8, c->d_dense, cd, d_changed);
                c->launches++;
            }
        }
        c->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        c->incremental_ok = false;
        return set_error(c, UVT_ERR_CUDA, "incremental commit failed: %s", cudaGetErrorString(e));
    }
    return UVT_OK;
}

// map_setVoxel (map.glsl:49-55): the write lands only where the chunk already holds a brick
int uvt_world_set_voxel(uvt_ctx *c, uint32_t x, uint32_t y, uint32_t z, uint32_t voxel, int *written) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    if (written) *written = 0;
    UVT_REQUIRE(c, c->h_chunks && c->h_bricks && c->world_committed, "no world committed");
    if (x >= c->dim || y >= c->dim || z >= c->dim) return UVT_OK;  // map_getChunkFlags reads 0 outside the map
    const uint32_t e = c->h_chunks[(size_t)(x >> 3) + (size_t)c->cd * ((size_t)(y >> 3) + (size_t)(z >> 3) * c->cd)];
    if (e == 0 || e > c->n_bricks) return UVT_OK;
    c->h_bricks[(size_t)(e - 1) * 512 + (x & 7u) + 8u * (y & 7u) + 64u * (z & 7u)] = voxel;
    if (written) *written = 1;
    const uint32_t p[3] = {x, y, z};
    return uvt_world_commit_region(c, c->n_bricks, p, p);
}

int uvt_world_layout_checksum(uvt_ctx *c, uint64_t out[4]) {
    if (!c || !out) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->world_committed, "no world committed");
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!c->compact_ok) return UVT_OK;  // reference layout only: nothing derived
    int rc = ensure_ready(c);            // material tables
    if (rc != UVT_OK) return rc;
    unsigned long long *d = nullptr;
    UVT_CUDA(c, cudaMalloc(&d, 32));
    cudaMemsetAsync(d, 0, 32, c->stream);
    const int cd = (int)c->cd;
    const int nq = (int)((c->dim / 4) * (c->dim / 4));
    layout_checksum_kernel<<<(unsigned)((size_t)cd * cd * cd), 64, 0, c->stream>>>(c->d_chunks2, c->d_bricks8, c->dense_valid ? c->d_dense : nullptr, c->d_mat_word, cd, d);
    clear4_checksum_kernel<<<(nq + 255) / 256, 256, 0, c->stream>>>(c->d_clear4, nq, d, 7ull);
    const int n1 = (int)(c->dim * c->dim);  // the sun clearance and its input follow the column tops as well
    clear4_checksum_kernel<<<(n1 + 255) / 256, 256, 0, c->stream>>>(c->d_top3, n1, d, 0xA3ull << 32);
    clear4_checksum_kernel<<<(n1 + 255) / 256, 256, 0, c->stream>>>(c->d_sun1, n1, d, 0x51ull << 32);
    c->launches += 4;
    unsigned long long h[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h, d, 32, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return set_error(c, UVT_ERR_CUDA, "layout checksum: %s", cudaGetErrorString(e));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
    out[3] = (uint64_t)(uint32_t)c->y_clear | ((uint64_t)(c->mat_words.size() - 1) << 32);
    return UVT_OK;
}

// ---- atlas --------------------------------------------------------------------------------------
int uvt_atlas_upload(uvt_ctx *c, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t w, uint32_t h, uint32_t d, const uint32_t *rgba) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, rgba, "rgba is NULL");
    UVT_REQUIRE(c, (uint64_t)ox + w <= 256 && (uint64_t)oy + h <= 256 && (uint64_t)oz + d <= 256, "sub-box leaves the 256^3 atlas");
    for (uint32_t z = 0; z < d; ++z)
        for (uint32_t y = 0; y < h; ++y)
            for (uint32_t x = 0; x < w; ++x) {
                const uint32_t ax = ox + x, ay = oy + y, az = oz + z;
                const uint32_t slot = (ax >> 3) + 32u * (ay >> 3) + 1024u * (az >> 3);
                if (slot >= c->n_slots) {
                    c->h_models.resize((size_t)(slot + 1) * 512, 0u);
                    c->n_slots = slot + 1;
                }
                c->h_models[(size_t)slot * 512 + (ax & 7u) + ((ay & 7u) << 3) + ((az & 7u) << 6)] = rgba[x + (size_t)w * (y + (size_t)h * z)];
            }
    c->atlas_dirty = true;
    return UVT_OK;
}

// ---- camera -------------------------------------------------------------------------------------
int uvt_set_camera(uvt_ctx *c, const uvt_camera *cam) { return uvt_set_cameras(c, cam, 1); }

int uvt_set_cameras(uvt_ctx *c, const uvt_camera *cams, int n) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, cams && n >= 1, "need at least one camera");
    c->cams.assign(cams, cams + n);
    derive_cam(cams[0], c->cam0);
    c->have_camera = true;
    if ((uint32_t)n != c->layers) {
        c->layers = (uint32_t)n;
        if (c->W && c->H) {
            UVT_CUDA(c, cudaStreamSynchronize(c->stream));
            int rc = alloc_gbuffer(c);
            if (rc != UVT_OK) return rc;
        }
    }
    if (n > 1) {
        if (n > c->d_cams_capacity) {
            UVT_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->d_cams);
            c->d_cams = nullptr;
            UVT_CUDA(c, cudaMalloc(&c->d_cams, sizeof(CamDev) * n));
            c->d_cams_capacity = n;
        }
        std::vector<CamDev> tmp(n);
        for (int i = 0; i < n; ++i) derive_cam(cams[i], tmp[i]);
        UVT_CUDA(c, cudaMemcpyAsync(c->d_cams, tmp.data(), sizeof(CamDev) * n, cudaMemcpyHostToDevice, c->stream));
        UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return UVT_OK;
}

// ---- G-buffer -----------------------------------------------------------------------------------
int uvt_resize(uvt_ctx *c, uint32_t width, uint32_t height) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, width >= 1 && height >= 1 && width <= 32768 && height <= 32768, "G-buffer size out of range");
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->W = width;
    c->H = height;
    return alloc_gbuffer(c);
}

int uvt_set_partition(uvt_ctx *c, uint32_t band_rows, uint32_t n_parts, uint32_t part) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, n_parts >= 1 && part < n_parts, "part must be < n_parts");
    UVT_REQUIRE(c, band_rows >= 8 && band_rows % 8 == 0 && band_rows % kTileH == 0, "band_rows must be a positive multiple of 8 (and of the CTA tile height)");
    UVT_REQUIRE(c, n_parts == 1 || band_rows % kPoolTile == 0, "with more than one part band_rows must be a multiple of 16 (the pooled CTA tile)");
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->band_rows = band_rows;
    c->n_parts = n_parts;
    c->part = part;
    if (c->W && c->H) return alloc_gbuffer(c);
    return UVT_OK;
}

int uvt_local_rows(uvt_ctx *c, uint32_t *rows) {
    if (!c || !rows) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    *rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
    return UVT_OK;
}

// ---- passes -------------------------------------------------------------------------------------
int uvt_dispatch_primary(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    int rc = pre_dispatch(c);
    if (rc != UVT_OK) return rc;
    PassTimer t(c, 0);
    return launch_primary<0>(c);
}

int uvt_dispatch_secondary(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    int rc = pre_dispatch(c);
    if (rc != UVT_OK) return rc;
    PassTimer t(c, 1);
    return launch_secondary<0>(c);
}

int uvt_dispatch_secondary_shade(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    int rc = pre_dispatch(c);
    if (rc != UVT_OK) return rc;
    if (!fuse_shade(c)) {  // pooled scheduler, or an entity pass sits between the two: the separate kernels
        rc = uvt_dispatch_secondary(c);
        return rc == UVT_OK ? uvt_shade(c) : rc;
    }
    PassTimer t(c, 1);
    return launch_secondary_shade(c, make_view(c, c->params.shadow_max_steps), make_gbuf(c), trace_grid(c));
}

int uvt_shade(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->W && c->H, "no G-buffer (uvt_resize first)");
    const uint32_t rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
    const dim3 grid((c->W + 63) / 64, (rows + 3) / 4, c->layers);
    PassTimer t(c, 2);
    shade_kernel<<<grid, 256, 0, c->stream>>>(make_view(c, 0), make_gbuf(c), make_target(c));
    return check_launch(c, "shade_kernel");
}

int uvt_dispatch_frame(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    int rc = pre_dispatch(c);
    if (rc != UVT_OK) return rc;
    const ViewDev v = make_view(c, c->params.primary_max_steps);
    const GBufDev g = make_gbuf(c);
    const dim3 grid = trace_grid(c);
    const CamDev *cams = c->layers > 1 ? c->d_cams : nullptr;
    PassTimer t(c, 3);
    const uint32_t all_rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
    if (c->frame_chunks > 1 && c->layers == 1 && !use_pool(c) && all_rows >= 32u * c->frame_chunks) {
        // Row chunks alternate between the ctx stream and a side stream: chunk k+1's primary pass fills the SMs that the tail of
        // chunk k's pass leaves idle (a pass ends with a few long-running warps; at 1/8 of a 4K frame per GPU that tail is a
        // quarter of the pass).  Same kernels, same pixels: results are identical to the whole-frame launches.
        ensure_sun(c);
        if (!c->side_stream) {
            UVT_CUDA(c, cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
            UVT_CUDA(c, cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
            UVT_CUDA(c, cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming));
        }
        c->ev_valid[0] = c->ev_valid[1] = c->ev_valid[2] = false;  // the passes of different chunks overlap: only the frame has a time
        UVT_CUDA(c, cudaEventRecord(c->fork_ev, c->stream));
        UVT_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->fork_ev, 0));
        cudaStream_t main_stream = c->stream;
        const uint32_t per = ((all_rows + c->frame_chunks - 1) / c->frame_chunks + 15u) & ~15u;  // multiples of the CTA tile height
        for (uint32_t k = 0, r0 = 0; r0 < all_rows && rc == UVT_OK; ++k, r0 += per) {
            c->stream = (k & 1u) ? c->side_stream : main_stream;
            rc = launch_rows(c, r0, std::min(per, all_rows - r0));
        }
        c->stream = main_stream;
        if (rc != UVT_OK) return rc;
        UVT_CUDA(c, cudaEventRecord(c->join_ev, c->side_stream));
        UVT_CUDA(c, cudaStreamWaitEvent(c->stream, c->join_ev, 0));
        return UVT_OK;
    }
    if (use_pool(c) || !(c->params.flags & UVT_FLAG_FUSED_FRAME) || ent_custom(c)) {
        // the three passes of game.zig:244-255 as three launches: measured faster than the fused kernel (c1 0.301 vs
        // 0.324 ms, c3 2.21 vs 2.41 ms) — the G-buffer round trip through L2 costs less than the registers and the
        // idle lanes of a kernel that keeps a primary and a shadow ray's state alive at once
        // (with timing on, every pass is bracketed by its own pair of events as well: uvt_last_pass_ms 0..2)
        {
            PassTimer tp(c, 0);
            rc = launch_primary<0>(c);
        }
        if (rc == UVT_OK && fuse_shade(c)) {  // shadow pass and blit in one launch; the blit's own timer brackets nothing
            {
                PassTimer ts(c, 1);
                rc = launch_secondary_shade(c, make_view(c, c->params.shadow_max_steps), g, grid);
            }
            PassTimer tb(c, 2);
            return rc;
        }
        if (rc == UVT_OK) {
            PassTimer ts(c, 1);
            rc = launch_secondary<0>(c);
        }
        if (rc != UVT_OK) return rc;
        const uint32_t rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part);
        PassTimer tb(c, 2);
        shade_kernel<<<dim3((c->W + 63) / 64, (rows + 3) / 4, c->layers), 256, 0, c->stream>>>(make_view(c, 0), g, make_target(c));
        return check_launch(c, "shade_kernel");
    }
    const bool batch = c->layers > 1;
    const uint32_t ss = c->params.shadow_max_steps;
    const FrameTarget ft = make_target(c);
    ensure_sun(c);
    if (use_dense(c)) {
        if (batch) frame_kernel<WorldDense, true, true><<<grid, kTileThreads, 0, c->stream>>>(world_dense(c), cams, c->cam0, v, ss, g, ft);
        else frame_kernel<WorldDense, true, false><<<grid, kTileThreads, 0, c->stream>>>(world_dense(c), cams, c->cam0, v, ss, g, ft);
    } else if (use_compact(c)) {
        if (batch) frame_kernel<WorldCompact, true, true><<<grid, kTileThreads, 0, c->stream>>>(world_compact(c), cams, c->cam0, v, ss, g, ft);
        else frame_kernel<WorldCompact, true, false><<<grid, kTileThreads, 0, c->stream>>>(world_compact(c), cams, c->cam0, v, ss, g, ft);
    } else {
        if (batch) frame_kernel<WorldRef, true, true><<<grid, kTileThreads, 0, c->stream>>>(world_ref(c), cams, c->cam0, v, ss, g, ft);
        else frame_kernel<WorldRef, true, false><<<grid, kTileThreads, 0, c->stream>>>(world_ref(c), cams, c->cam0, v, ss, g, ft);
    }
    return check_launch(c, "frame_kernel");
}

// ---- entities -----------------------------------------------------------------------------------
int uvt_set_entity_mode(uvt_ctx *c, uint32_t mode) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, mode == UVT_ENTITY_BOXES || mode == UVT_ENTITY_MODELS, "entity mode out of range");
    c->ent_mode = mode;
    if (mode == UVT_ENTITY_MODELS && !(c->params.flags & UVT_FLAG_HIT_BUFFER)) {
        // the composite reads distance(C_position, inter.hit_pos / 8) (primary.comp.glsl:47) from the hit buffer
        c->params.flags |= UVT_FLAG_HIT_BUFFER;
        if (c->W && c->H) {
            UVT_CUDA(c, cudaStreamSynchronize(c->stream));
            return alloc_gbuffer(c);
        }
    }
    return UVT_OK;
}

int uvt_set_entities(uvt_ctx *c, const float *positions_xyz, uint32_t n) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, n <= (uint32_t)kMaxEntities, "at most 32 entities");
    UVT_REQUIRE(c, n == 0 || positions_xyz, "positions is NULL");
    if (n == 0) c->ent_pos.clear();  // back to the literal positions of map.glsl:173-179
    else c->ent_pos.assign(positions_xyz, positions_xyz + 3 * (size_t)n);
    return UVT_OK;
}

int uvt_entity_model_upload(uvt_ctx *c, uint32_t size, const uint32_t *rgba, uint32_t max_steps) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    const uint32_t steps = max_steps ? max_steps : 64u;
    UVT_REQUIRE(c, steps <= 65535u, "entity step cap above 65535");
    UVT_REQUIRE(c, !rgba || size == 8 || size == 16 || size == 32, "entity model edge must be 8, 16 or 32 voxels");
    uint32_t *d_new = nullptr;
    if (rgba) {  // build the new model first: a failure leaves the current one in place
        const size_t bytes = (size_t)size * size * size * 4;
        UVT_CUDA(c, cudaMalloc(&d_new, bytes));
        cudaError_t e = cudaMemcpy(d_new, rgba, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d_new);
            return set_error(c, UVT_ERR_CUDA, "entity model upload: %s", cudaGetErrorString(e));
        }
    }
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));  // a frame in flight may still read the old model
    cudaFree(c->d_ent_model);
    c->d_ent_model = d_new;          // nullptr: back to the atlas texels [0,8)^3
    c->ent_size = rgba ? size : 8u;
    c->ent_steps = steps;
    return UVT_OK;
}

int uvt_set_frame_chunks(uvt_ctx *c, uint32_t n) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, n >= 1 && n <= 8, "1..8 row chunks");
    c->frame_chunks = n;
    return UVT_OK;
}

int uvt_pick(uvt_ctx *c, uvt_hit *out) {
    if (!c || !out) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->have_camera, "no camera (uvt_set_camera first)");
    int rc = ensure_ready(c);
    if (rc != UVT_OK) return rc;
    ViewDev v = make_view(c, c->params.edit_max_steps);
    if (use_compact(c)) pick_kernel<WorldCompact><<<1, 128, 0, c->stream>>>(world_compact(c), c->cam0, v, c->d_pick);
    else pick_kernel<WorldRef><<<1, 128, 0, c->stream>>>(world_ref(c), c->cam0, v, c->d_pick);
    rc = check_launch(c, "pick_kernel");
    if (rc != UVT_OK) return rc;
    UVT_CUDA(c, cudaMemcpyAsync(out, c->d_pick, sizeof(uvt_hit), cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    return UVT_OK;
}

int uvt_sync(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    return UVT_OK;
}

static int buffer_info(uvt_ctx *c, uvt_buffer_kind kind, void **ptr, size_t *bytes_per_px) {
    switch (kind) {
        case UVT_BUF_ALBEDO: *ptr = c->d_albedo; *bytes_per_px = 4; break;
        case UVT_BUF_NORMAL: *ptr = c->d_normal; *bytes_per_px = 4; break;
        case UVT_BUF_POSITION: *ptr = c->d_position; *bytes_per_px = 16; break;
        case UVT_BUF_ILLUMINATION: *ptr = c->d_illum; *bytes_per_px = 4; break;
        case UVT_BUF_FRAME: *ptr = c->d_frame; *bytes_per_px = 4; break;
        case UVT_BUF_HIT: *ptr = c->d_hit; *bytes_per_px = sizeof(uvt_hit); break;
        default: return set_error(c, UVT_ERR_INVALID, "unknown buffer kind %d", (int)kind);
    }
    if (!*ptr) return set_error(c, UVT_ERR_INVALID, "buffer %d is not allocated (uvt_resize / UVT_FLAG_HIT_BUFFER)", (int)kind);
    return UVT_OK;
}

size_t uvt_buffer_bytes(uvt_ctx *c, uvt_buffer_kind kind) {
    if (!c) return 0;
    UVT_ENTER(c);
    void *p = nullptr;
    size_t bpp = 0;
    if (buffer_info(c, kind, &p, &bpp) != UVT_OK) return 0;
    return c->gbuf_pixels * bpp;
}

int uvt_readback(uvt_ctx *c, uvt_buffer_kind kind, void *dst, size_t bytes) {
    if (!c || !dst) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    void *p = nullptr;
    size_t bpp = 0;
    int rc = buffer_info(c, kind, &p, &bpp);
    if (rc != UVT_OK) return rc;
    UVT_REQUIRE(c, bytes <= c->gbuf_pixels * bpp, "readback larger than the buffer");
    UVT_CUDA(c, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    return UVT_OK;
}

int uvt_readback_async(uvt_ctx *c, uvt_buffer_kind kind, void *dst, size_t bytes) {
    if (!c || !dst) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    void *p = nullptr;
    size_t bpp = 0;
    int rc = buffer_info(c, kind, &p, &bpp);
    if (rc != UVT_OK) return rc;
    UVT_REQUIRE(c, bytes <= c->gbuf_pixels * bpp, "readback larger than the buffer");
    const int s = c->snap_next;
    c->snap_next ^= 1;
    if (c->snap_busy[s]) {  // the snapshot slot is reused: its previous copy must have landed
        UVT_CUDA(c, cudaEventSynchronize(c->snap_done[s]));
        c->snap_busy[s] = false;
    }
    if (c->snap_bytes[s] < bytes) {
        cudaFree(c->snap[s]);
        c->snap[s] = nullptr;
        c->snap_bytes[s] = 0;
        UVT_CUDA(c, cudaMalloc(&c->snap[s], bytes));
        c->snap_bytes[s] = bytes;
    }
    UVT_CUDA(c, cudaMemcpyAsync(c->snap[s], p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    UVT_CUDA(c, cudaEventRecord(c->snap_ready[s], c->stream));
    UVT_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->snap_ready[s], 0));
    UVT_CUDA(c, cudaMemcpyAsync(dst, c->snap[s], bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    UVT_CUDA(c, cudaEventRecord(c->snap_done[s], c->copy_stream));
    c->snap_busy[s] = true;
    return UVT_OK;
}

// The band-partitioned FRAME of this ctx copied straight to where its bands belong in a full W x H host frame: every
// rank of a tiled frame calls this on the SAME host frame (shared memory registered with uvt_host_register), so the
// frame is assembled in host memory by N device-to-host copies running in parallel on N PCIe links — no GPU-to-GPU
// exchange, no funnel through the presenting GPU.  Pipelined like uvt_readback_async (snapshot, copy stream).
int uvt_readback_bands_async(uvt_ctx *c, void *host_frame, size_t frame_bytes) {
    if (!c || !host_frame) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->d_frame && c->layers == 1, "no single-layer frame buffer (uvt_resize first)");
    UVT_REQUIRE(c, frame_bytes >= (size_t)c->W * c->H * 4, "host frame smaller than W*H*4 bytes");
    const size_t row = (size_t)c->W * 4, rows = storage_rows(c->H, c->band_rows, c->n_parts, c->part), bytes = rows * row;
    if (bytes == 0) return UVT_OK;
    const int s = c->snap_next;
    c->snap_next ^= 1;
    if (c->snap_busy[s]) {
        UVT_CUDA(c, cudaEventSynchronize(c->snap_done[s]));
        c->snap_busy[s] = false;
    }
    if (c->snap_bytes[s] < bytes) {
        cudaFree(c->snap[s]);
        c->snap[s] = nullptr;
        c->snap_bytes[s] = 0;
        UVT_CUDA(c, cudaMalloc(&c->snap[s], bytes));
        c->snap_bytes[s] = bytes;
    }
    UVT_CUDA(c, cudaMemcpyAsync(c->snap[s], c->d_frame, bytes, cudaMemcpyDeviceToDevice, c->stream));
    UVT_CUDA(c, cudaEventRecord(c->snap_ready[s], c->stream));
    UVT_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->snap_ready[s], 0));
    if (c->n_parts == 1) {
        UVT_CUDA(c, cudaMemcpyAsync(host_frame, c->snap[s], (size_t)c->H * row, cudaMemcpyDeviceToHost, c->copy_stream));
    } else {
        // local band lb is global band lb * n_parts + part; all but possibly the frame's last band are full
        const size_t band = (size_t)c->band_rows * row;
        const uint32_t n_local = (uint32_t)(rows / c->band_rows);
        const uint32_t last_global = (n_local - 1) * c->n_parts + c->part;
        const uint32_t last_rows = std::min<uint32_t>(c->band_rows, c->H - last_global * c->band_rows);
        const uint32_t n_full = last_rows == c->band_rows ? n_local : n_local - 1;
        char *dst = (char *)host_frame + (size_t)c->part * band;
        if (n_full)
            UVT_CUDA(c, cudaMemcpy2DAsync(dst, band * c->n_parts, c->snap[s], band, band, n_full, cudaMemcpyDeviceToHost, c->copy_stream));
        if (n_full != n_local)
            UVT_CUDA(c, cudaMemcpyAsync(dst + (size_t)n_full * band * c->n_parts, (char *)c->snap[s] + (size_t)n_full * band,
                                        (size_t)last_rows * row, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    UVT_CUDA(c, cudaEventRecord(c->snap_done[s], c->copy_stream));
    c->snap_busy[s] = true;
    return UVT_OK;
}

// Page-lock caller memory (e.g. a POSIX shared-memory frame mapped by every rank) so that device-to-host copies into it
// run asynchronously at full PCIe rate.
int uvt_host_register(uvt_ctx *c, void *p, size_t bytes) {
    if (!c || !p || !bytes) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_CUDA(c, cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return UVT_OK;
}

int uvt_host_unregister(uvt_ctx *c, void *p) {
    if (!c || !p) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_CUDA(c, cudaHostUnregister(p));
    return UVT_OK;
}

int uvt_readback_wait(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    for (int s = 0; s < 2; ++s)
        if (c->snap_busy[s]) {
            UVT_CUDA(c, cudaEventSynchronize(c->snap_done[s]));
            c->snap_busy[s] = false;
        }
    return UVT_OK;
}

int uvt_device_ptr(uvt_ctx *c, uvt_buffer_kind kind, void **dptr) {
    if (!c || !dptr) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    size_t bpp;
    return buffer_info(c, kind, dptr, &bpp);
}

int uvt_bind_frame_target(uvt_ctx *c, void *dptr, uint32_t row_offset_rows, uint32_t global_rows) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    (void)row_offset_rows;
    c->frame_target = (uint32_t *)dptr;
    c->frame_target_global_rows = global_rows != 0;
    return UVT_OK;
}

int uvt_shared_frame_create(uvt_ctx *c, void **dptr, unsigned char handle_out[64]) {
    if (!c || !dptr || !handle_out) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->W && c->H, "no G-buffer (uvt_resize first)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->shared_frame);
    c->shared_frame = nullptr;
    UVT_CUDA(c, cudaMalloc(&c->shared_frame, (size_t)c->W * c->H * 4));
    UVT_CUDA(c, cudaMemset(c->shared_frame, 0, (size_t)c->W * c->H * 4));
    cudaIpcMemHandle_t h;
    UVT_CUDA(c, cudaIpcGetMemHandle(&h, c->shared_frame));
    std::memcpy(handle_out, &h, 64);
    *dptr = c->shared_frame;
    return UVT_OK;
}

int uvt_shared_frame_open(uvt_ctx *c, const unsigned char handle[64], void **dptr) {
    if (!c || !dptr || !handle) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    UVT_CUDA(c, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return UVT_OK;
}

int uvt_shared_frame_close(uvt_ctx *c, void *dptr) {
    if (!c || !dptr) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->frame_target == dptr) c->frame_target = nullptr;
    UVT_CUDA(c, cudaIpcCloseMemHandle(dptr));
    return UVT_OK;
}

int uvt_read_device(uvt_ctx *c, const void *dptr, void *dst, size_t bytes) {
    if (!c || !dptr || !dst) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_CUDA(c, cudaMemcpyAsync(dst, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    return UVT_OK;
}

int uvt_alloc_pinned(uvt_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_CUDA(c, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    c->pinned.push_back(*out);  // owned by the ctx: whatever the caller has not freed goes with uvt_destroy
    return UVT_OK;
}

int uvt_free_pinned(uvt_ctx *c, void *p) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    if (!p) return UVT_OK;
    auto it = std::find(c->pinned.begin(), c->pinned.end(), p);
    UVT_REQUIRE(c, it != c->pinned.end(), "not a live uvt_alloc_pinned allocation of this ctx");
    c->pinned.erase(it);
    UVT_CUDA(c, cudaFreeHost(p));
    return UVT_OK;
}

int uvt_count_pass(uvt_ctx *c, int which, uvt_counters *out) {
    if (!c || !out) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, which >= 0 && which <= 3, "which must be 0/1 (reference counters) or 2/3 (fast-path fetch statistics) for primary/secondary");
    int rc = pre_dispatch(c);
    if (rc != UVT_OK) return rc;
    UVT_CUDA(c, cudaMemsetAsync(c->d_counters, 0, sizeof(DevCounters), c->stream));
    rc = which == 0 ? launch_primary<1>(c) : (which == 1 ? launch_secondary<1>(c) : (which == 2 ? launch_primary<2>(c) : launch_secondary<2>(c)));
    if (rc != UVT_OK) return rc;
    DevCounters h;
    UVT_CUDA(c, cudaMemcpyAsync(&h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    out->rays = h.rays; out->t_in = h.t_in; out->t_chunk = h.t_chunk; out->t_block = h.t_block;
    out->hits = h.hits; out->early_out = h.early_out;
    return UVT_OK;
}

int uvt_enable_timing(uvt_ctx *c, int on) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    c->timing = on != 0;
    return UVT_OK;
}

int uvt_last_pass_ms(uvt_ctx *c, int which, float *ms) {
    if (!c || !ms) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, which >= 0 && which < 4, "which must be 0..3");
    UVT_REQUIRE(c, c->ev_valid[which], "pass has not run with timing enabled");
    UVT_CUDA(c, cudaEventSynchronize(c->ev[which][1]));
    UVT_CUDA(c, cudaEventElapsedTime(ms, c->ev[which][0], c->ev[which][1]));
    return UVT_OK;
}

uint64_t uvt_launch_count(uvt_ctx *c) { return c ? c->launches : 0; }

int uvt_deinterleave(uvt_ctx *c, const void *gathered, void *frame, uint32_t rows_per_part) {
    if (!c || !gathered || !frame) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    const dim3 grid((c->W + 255) / 256, c->H, 1);
    deinterleave_kernel<<<grid, 256, 0, c->stream>>>((const uint32_t *)gathered, (uint32_t *)frame, c->W, c->H, c->band_rows, c->n_parts, rows_per_part);
    return check_launch(c, "deinterleave_kernel");
}

int uvt_measure_l2_read_gbps(uvt_ctx *c, size_t bytes, int repeats, float *gbps) {
    if (!c || !gbps) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, bytes >= 4096 && bytes % 16 == 0 && repeats >= 1, "bytes must be a multiple of 16, repeats >= 1");
    uint4 *buf = nullptr;
    UVT_CUDA(c, cudaMalloc(&buf, bytes));
    UVT_CUDA(c, cudaMemsetAsync(buf, 1, bytes, c->stream));
    const int blocks = c->sm_count * 8;
    l2_read_kernel<<<blocks, 256, 0, c->stream>>>(buf, bytes / 16, 2, c->d_sink);  // warm: pull the buffer into L2
    c->launches++;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, c->stream);
    l2_read_kernel<<<blocks, 256, 0, c->stream>>>(buf, bytes / 16, repeats, c->d_sink);
    cudaEventRecord(b, c->stream);
    int rc = check_launch(c, "l2_read_kernel");
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf);
    if (rc != UVT_OK) return rc;
    *gbps = (float)((double)bytes * repeats / (ms * 1e-3) / 1e9);
    return UVT_OK;
}

int uvt_measure_hbm_copy_gbps(uvt_ctx *c, size_t bytes, int repeats, float *gbps) {
    if (!c || !gbps) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, bytes >= 4096 && bytes % 16 == 0 && repeats >= 1, "bytes must be a multiple of 16, repeats >= 1");
    uint4 *src = nullptr, *dst = nullptr;
    UVT_CUDA(c, cudaMalloc(&src, bytes));
    UVT_CUDA(c, cudaMalloc(&dst, bytes));
    UVT_CUDA(c, cudaMemsetAsync(src, 1, bytes, c->stream));
    const int blocks = c->sm_count * 16;
    copy_kernel<<<blocks, 256, 0, c->stream>>>(src, dst, bytes / 16);
    c->launches++;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, c->stream);
    for (int r = 0; r < repeats; ++r) {
        copy_kernel<<<blocks, 256, 0, c->stream>>>(src, dst, bytes / 16);
        c->launches++;
    }
    cudaEventRecord(b, c->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(src); cudaFree(dst);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(c, UVT_ERR_CUDA, "copy_kernel: %s", cudaGetErrorString(e));
    *gbps = (float)(2.0 * (double)bytes * repeats / (ms * 1e-3) / 1e9);
    return UVT_OK;
}

}  // extern "C"

// ---- NCCL band exchange ---------------------------------------------------------------
// The frame is cut into interleaved bands (uvt_set_partition); with NCCL every non-presenting rank SENDS its finished
// bands and rank 0 RECEIVES each of them straight at its place in the full frame (a band is a contiguous run of
// band_rows * W pixels there), so no reassembly pass is needed.  NCCL has no gather: it is grouped ncclSend / ncclRecv.
// The rank's bands are rendered in band GROUPS; the exchange of group g runs on a second stream while group g + 1 is
// traversed.  libnccl is bound at run time: in a torchrun process that is the copy torch already loaded.
namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    if (api.lib || !api.error.empty()) return &api;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.error = std::string("libnccl.so.2 cannot be loaded: ") + dlerror();
        return &api;
    }
#define UVT_NCCL_SYM(field, sym)                                                        \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym));              \
    if (!api.field && api.error.empty()) api.error = std::string("libnccl lacks ") + sym;
    UVT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    UVT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    UVT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    UVT_NCCL_SYM(GroupStart, "ncclGroupStart")
    UVT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    UVT_NCCL_SYM(Send, "ncclSend")
    UVT_NCCL_SYM(Recv, "ncclRecv")
    UVT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef UVT_NCCL_SYM
    return &api;
}

#define UVT_NCCL(ctx, expr)                                                                                   \
    do {                                                                                                      \
        ncclResult_t r_ = (expr);                                                                             \
        if (r_ != ncclSuccess) return set_error((ctx), UVT_ERR_CUDA, "%s: %s", #expr, nccl_api()->GetErrorString(r_)); \
    } while (0)

// rows [row0, row0 + nrows) of the ctx's local storage through the three passes of the frame (tile scheduler)
int launch_rows(uvt_ctx *c, uint32_t row0, uint32_t nrows) {
    ViewDev v = make_view(c, c->params.primary_max_steps);
    v.row0 = row0;
    const GBufDev g = make_gbuf(c);
    const dim3 grid((c->W + kTileW - 1) / kTileW, (nrows + kTileH - 1) / kTileH, 1);
    if (use_dense(c)) launch_primary_world<WorldDense, 0>(c, world_dense(c), v, g, grid);
    else if (use_compact(c)) launch_primary_world<WorldCompact, 0>(c, world_compact(c), v, g, grid);
    else launch_primary_world<WorldRef, 0>(c, world_ref(c), v, g, grid);
    int rc = check_launch(c, "primary_kernel");
    if (rc == UVT_OK) rc = launch_entity_primary(c, v, g, nrows, 1);
    if (rc != UVT_OK) return rc;
    v.max_steps = c->params.shadow_max_steps;
    if (fuse_shade(c)) return launch_secondary_shade(c, v, g, grid);
    if (use_dense(c)) secondary_kernel<WorldDense, 0><<<grid, kTileThreads, 0, c->stream>>>(world_dense(c), v, g, c->d_counters);
    else if (use_compact(c)) secondary_kernel<WorldCompact, 0><<<grid, kTileThreads, 0, c->stream>>>(world_compact(c), v, g, c->d_counters);
    else secondary_kernel<WorldRef, 0><<<grid, kTileThreads, 0, c->stream>>>(world_ref(c), v, g, c->d_counters);
    rc = check_launch(c, "secondary_kernel");
    if (rc == UVT_OK) rc = launch_entity_shadow(c, v, g, nrows, 1);
    if (rc != UVT_OK) return rc;
    shade_kernel<<<dim3((c->W + 63) / 64, (nrows + 3) / 4, 1), 256, 0, c->stream>>>(v, g, make_target(c));
    return check_launch(c, "shade_kernel");
}

}  // namespace

extern "C" {

int uvt_nccl_unique_id(unsigned char id_out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NcclApi *api = nccl_api();
    if (!id_out) return UVT_ERR_INVALID;
    if (!api->error.empty()) return set_error(nullptr, UVT_ERR_INVALID, "%s", api->error.c_str());
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return set_error(nullptr, UVT_ERR_CUDA, "ncclGetUniqueId: %s", api->GetErrorString(r));
    std::memcpy(id_out, &id, 128);
    return UVT_OK;
}

int uvt_nccl_init(uvt_ctx *c, const unsigned char id[128], int n_ranks, int rank) {
    if (!c || !id) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    NcclApi *api = nccl_api();
    UVT_REQUIRE(c, api->error.empty(), api->error.c_str());
    UVT_REQUIRE(c, n_ranks >= 1 && rank >= 0 && rank < n_ranks, "rank out of range");
    UVT_REQUIRE(c, !c->nccl, "uvt_nccl_init called twice (uvt_nccl_shutdown first)");
    ncclUniqueId uid;
    std::memcpy(&uid, id, 128);
    UVT_NCCL(c, api->CommInitRank(&c->nccl, n_ranks, uid, rank));
    c->nccl_ranks = n_ranks;
    c->nccl_rank = rank;
    {   // the exchange kernels must not queue behind the CTAs of the next band group: highest stream priority
        int lo = 0, hi = 0;
        UVT_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        UVT_CUDA(c, cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    }
    for (auto &e : c->group_done) UVT_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    UVT_CUDA(c, cudaEventCreateWithFlags(&c->exchange_done, cudaEventDisableTiming));
    return UVT_OK;
}

int uvt_nccl_shutdown(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    if (!c->nccl) return UVT_OK;
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->comm_stream);
    nccl_api()->CommDestroy(c->nccl);
    c->nccl = nullptr;
    cudaStreamDestroy(c->comm_stream);
    c->comm_stream = nullptr;
    for (auto &e : c->group_done) { cudaEventDestroy(e); e = nullptr; }
    cudaEventDestroy(c->exchange_done);
    c->exchange_done = nullptr;
    return UVT_OK;
}

int uvt_dispatch_frame_nccl(uvt_ctx *c, void *full_frame, uint32_t n_groups) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    NcclApi *api = nccl_api();
    UVT_REQUIRE(c, c->nccl, "no NCCL communicator (uvt_nccl_init first)");
    UVT_REQUIRE(c, c->n_parts == (uint32_t)c->nccl_ranks && c->part == (uint32_t)c->nccl_rank, "uvt_set_partition(band, n_ranks, rank) must match the communicator");
    UVT_REQUIRE(c, c->layers == 1, "batched poses are not tiled");
    UVT_REQUIRE(c, n_groups >= 1 && n_groups <= 8, "1..8 band groups");
    UVT_REQUIRE(c, c->nccl_rank != 0 || full_frame, "the presenting rank passes the W x H frame to assemble");
    int rc = pre_dispatch(c);
    if (rc != UVT_OK) return rc;
    ensure_sun(c);
    const uint32_t band = c->band_rows, n = c->n_parts, H = c->H;
    const size_t band_px = (size_t)band * c->W;
    const uint32_t n_bands = (H + band - 1) / band;
    auto local_bands = [&](uint32_t part) { return part < n_bands ? (n_bands - part + n - 1) / n : 0u; };  // bands part, part + n, ...
    const uint32_t mine = local_bands(c->part);
    const uint32_t most = local_bands(0);                      // the largest part: every rank walks the same group boundaries
    const uint32_t per_group = (most + n_groups - 1) / n_groups;
    // the presenting rank's own pixels go straight to their place; the others shade into their compact band buffer
    uint32_t *saved_target = c->frame_target;
    const bool saved_global = c->frame_target_global_rows;
    if (c->nccl_rank == 0) { c->frame_target = (uint32_t *)full_frame; c->frame_target_global_rows = true; }
    else { c->frame_target = nullptr; c->frame_target_global_rows = false; }
    PassTimer t(c, 3);
    for (uint32_t g = 0; g < n_groups && rc == UVT_OK; ++g) {
        const uint32_t lb0 = g * per_group, lb1 = std::min((g + 1) * per_group, most);
        if (lb0 >= lb1) break;
        if (lb0 < mine) rc = launch_rows(c, lb0 * band, (std::min(lb1, mine) - lb0) * band);
        if (rc != UVT_OK) break;
        if (n == 1) continue;
        UVT_CUDA(c, cudaEventRecord(c->group_done[g], c->stream));
        UVT_CUDA(c, cudaStreamWaitEvent(c->comm_stream, c->group_done[g], 0));
        UVT_NCCL(c, api->GroupStart());
        if (c->nccl_rank == 0) {
            for (uint32_t p = 1; p < n; ++p)
                for (uint32_t lb = lb0; lb < std::min(lb1, local_bands(p)); ++lb) {
                    const uint32_t gb = lb * n + p;  // global band
                    const size_t px = (size_t)std::min(band, H - gb * band) * c->W;
                    UVT_NCCL(c, api->Recv((uint32_t *)full_frame + (size_t)gb * band_px, px, ncclUint32, (int)p, c->nccl, c->comm_stream));
                }
        } else {
            for (uint32_t lb = lb0; lb < std::min(lb1, mine); ++lb) {
                const uint32_t gb = lb * n + c->part;
                const size_t px = (size_t)std::min(band, H - gb * band) * c->W;
                UVT_NCCL(c, api->Send(c->d_frame + (size_t)lb * band_px, px, ncclUint32, 0, c->nccl, c->comm_stream));
            }
        }
        UVT_NCCL(c, api->GroupEnd());
    }
    c->frame_target = saved_target;
    c->frame_target_global_rows = saved_global;
    if (rc != UVT_OK) return rc;
    if (n > 1) {  // the frame is complete on the ctx stream once the last exchange has landed
        UVT_CUDA(c, cudaEventRecord(c->exchange_done, c->comm_stream));
        UVT_CUDA(c, cudaStreamWaitEvent(c->stream, c->exchange_done, 0));
    }
    return UVT_OK;
}

}  // extern "C"

// ---- procgen on the device (src/procgen.zig:6-70) --------------------------------------
extern "C" {

static void procgen_release(uvt_ctx *c) {
    cudaFree(c->pgen.vh); cudaFree(c->pgen.seeds); cudaFree(c->pgen.deco); cudaFree(c->pgen.trees);
    c->pgen.vh = nullptr; c->pgen.seeds = c->pgen.deco = nullptr; c->pgen.trees = nullptr;
    c->pgen.planned = false;
}

int uvt_world_procgen_plan(uvt_ctx *c, float offset_x, float offset_y, size_t *n_bricks) {
    if (!c || !n_bricks) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->h_chunks && c->d_chunks, "no world allocated (uvt_world_alloc first)");
    UVT_REQUIRE(c, c->dim >= 16, "procgen needs a map of at least 16 blocks (the y < 16 slab)");
    namespace pg = uvt::pg;
    procgen_release(c);
    const uint32_t dim = c->dim, cd = c->cd;
    const size_t n_col = (size_t)dim * dim, n_chunks = (size_t)cd * cd * cd;
    const uint32_t max_trees = 1u << 16;

    // LCG jump table and gradient table (host, tiny)
    std::vector<uint2> jump(2 * (size_t)dim + 64);
    jump[0] = make_uint2(1u, 0u);
    for (size_t n = 1; n < jump.size(); ++n) jump[n] = make_uint2(jump[n - 1].x * pg::kLcgA, jump[n - 1].y * pg::kLcgA + pg::kLcgC);
    float grad[256];
    uvt_noise::build_grad_table(grad);

    uint2 *d_jump = nullptr;
    float *d_grad = nullptr;
    ushort2 *d_blocked = nullptr;
    uint8_t *d_nblocked = nullptr, *d_skipped = nullptr;
    uint32_t *d_misc = nullptr;  // [0] n_trees [1] status [2] final seed [3] n_touched
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_idx = nullptr, *d_idx2 = nullptr;
    void *d_tmp = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_jump); cudaFree(d_grad); cudaFree(d_blocked); cudaFree(d_nblocked); cudaFree(d_skipped); cudaFree(d_misc);
        cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_idx); cudaFree(d_idx2); cudaFree(d_tmp);
    };
#define UVT_PG(expr)                                                                                                    \
    do {                                                                                                                \
        cudaError_t e_ = (expr);                                                                                        \
        if (e_ != cudaSuccess) {                                                                                        \
            cleanup();                                                                                                  \
            procgen_release(c);                                                                                         \
            return set_error(c, e_ == cudaErrorMemoryAllocation ? UVT_ERR_OOM : UVT_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
        }                                                                                                               \
    } while (0)
    UVT_PG(cudaMalloc(&c->pgen.vh, n_col * 2));
    UVT_PG(cudaMalloc(&c->pgen.seeds, n_col * 4));
    UVT_PG(cudaMalloc(&c->pgen.deco, n_col * 4));
    UVT_PG(cudaMalloc(&c->pgen.trees, (size_t)max_trees * sizeof(pg::Tree)));
    UVT_PG(cudaMalloc(&d_jump, jump.size() * sizeof(uint2)));
    UVT_PG(cudaMalloc(&d_grad, sizeof grad));
    UVT_PG(cudaMalloc(&d_blocked, (size_t)3 * dim * pg::kMaxRanges * sizeof(ushort2)));
    UVT_PG(cudaMalloc(&d_nblocked, (size_t)3 * dim));
    UVT_PG(cudaMalloc(&d_skipped, n_col));
    UVT_PG(cudaMalloc(&d_misc, 16));
    UVT_PG(cudaMalloc(&d_keys, n_chunks * 8));
    UVT_PG(cudaMalloc(&d_keys2, n_chunks * 8));
    UVT_PG(cudaMalloc(&d_idx, n_chunks * 4));
    UVT_PG(cudaMalloc(&d_idx2, n_chunks * 4));
    UVT_PG(cudaMemcpyAsync(d_jump, jump.data(), jump.size() * sizeof(uint2), cudaMemcpyHostToDevice, c->stream));
    UVT_PG(cudaMemcpyAsync(d_grad, grad, sizeof grad, cudaMemcpyHostToDevice, c->stream));
    UVT_PG(cudaMemsetAsync(d_skipped, 0, n_col, c->stream));
    UVT_PG(cudaMemsetAsync(d_nblocked, 0, (size_t)3 * dim, c->stream));
    UVT_PG(cudaMemsetAsync(d_misc, 0, 16, c->stream));
    UVT_PG(cudaMemsetAsync(d_keys, 0xFF, n_chunks * 8, c->stream));
    UVT_PG(cudaMemsetAsync(c->d_chunks, 0, n_chunks * 4, c->stream));

    pg::heights_kernel<<<dim3((dim + 127) / 128, dim), 128, 0, c->stream>>>(d_grad, dim, offset_x, offset_y, c->pgen.vh);
    pg::scan_kernel<<<1, pg::kScanThreads, 0, c->stream>>>(c->pgen.vh, dim, d_jump, 0x46AE4Fu, c->pgen.seeds, d_skipped, c->pgen.trees, max_trees, d_misc,
                                                           d_blocked, d_nblocked, d_misc + 1, d_misc + 2);
    pg::deco_kernel<<<(unsigned)((n_col + 255) / 256), 256, 0, c->stream>>>(c->pgen.vh, c->pgen.seeds, d_skipped, d_jump, n_col, c->pgen.deco);
    c->launches += 3;
    uint32_t misc[4] = {0, 0, 0, 0};
    UVT_PG(cudaMemcpyAsync(misc, d_misc, 16, cudaMemcpyDeviceToHost, c->stream));
    UVT_PG(cudaStreamSynchronize(c->stream));
    if (misc[1] != 0) {  // more tree ranges / trees than the device path keeps: the caller falls back to the serial host procgen
        cleanup();
        procgen_release(c);
        return set_error(c, UVT_ERR_INVALID, "device procgen: unsupported world (status %u)", misc[1]);
    }
    c->pgen.n_trees = misc[0];

    pg::touch_kernel<<<(cd * cd + 127) / 128, 128, 0, c->stream>>>(c->pgen.vh, c->pgen.deco, dim, d_keys);
    if (c->pgen.n_trees) pg::tree_touch_kernel<<<(c->pgen.n_trees + 63) / 64, 64, 0, c->stream>>>(c->pgen.trees, c->pgen.n_trees, dim, d_keys, d_misc + 1);
    pg::iota_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, c->stream>>>(d_idx, (uint32_t)n_chunks);
    size_t tmp_bytes = 0;
    UVT_PG(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)n_chunks, 0, 64, c->stream));
    UVT_PG(cudaMalloc(&d_tmp, tmp_bytes));
    UVT_PG(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)n_chunks, 0, 64, c->stream));
    pg::slab_chunks_kernel<<<(2 * cd * cd + 255) / 256, 256, 0, c->stream>>>(c->d_chunks, cd);
    pg::ranked_chunks_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, c->stream>>>(d_keys2, d_idx2, (uint32_t)n_chunks, 2 * cd * cd, c->d_chunks, d_misc + 3);
    c->launches += 5;
    UVT_PG(cudaMemcpyAsync(misc, d_misc, 16, cudaMemcpyDeviceToHost, c->stream));
    UVT_PG(cudaStreamSynchronize(c->stream));
    UVT_PG(cudaGetLastError());
#undef UVT_PG
    cleanup();
    if (misc[1] != 0) {
        procgen_release(c);
        return set_error(c, UVT_ERR_INVALID, "device procgen: unsupported world (status %u)", misc[1]);
    }
    c->pgen.n_bricks = (size_t)2 * cd * cd + misc[3];
    c->pgen.planned = true;
    *n_bricks = c->pgen.n_bricks;
    return UVT_OK;
}

int uvt_world_procgen_fill(uvt_ctx *c) {
    if (!c) return UVT_ERR_INVALID;
    UVT_ENTER(c);
    UVT_REQUIRE(c, c->pgen.planned, "uvt_world_procgen_plan first");
    UVT_REQUIRE(c, c->pgen.n_bricks <= c->h_capacity, "the staging holds fewer bricks than the world needs (uvt_world_grow first)");
    namespace pg = uvt::pg;
    const uint32_t dim = c->dim, cd = c->cd;
    const size_t n = c->pgen.n_bricks, n_chunks = (size_t)cd * cd * cd;
    if (n > c->d_brick_capacity) {
        cudaFree(c->d_bricks);
        c->d_bricks = nullptr;
        c->d_brick_capacity = 0;
        const size_t want = std::max(n, c->h_capacity);
        UVT_CUDA(c, cudaMalloc(&c->d_bricks, want * 2048));
        c->d_brick_capacity = want;
    }
    UVT_CUDA(c, cudaMemsetAsync(c->d_bricks, 0, n * 2048, c->stream));
    pg::fill_kernel<<<dim3((dim + 63) / 64, dim), 64, 0, c->stream>>>(c->pgen.vh, c->pgen.seeds, c->pgen.deco, c->d_chunks, dim, c->d_bricks);
    pg::tree_fill_kernel<<<1, 32, 0, c->stream>>>(c->pgen.trees, c->pgen.n_trees, c->pgen.vh, c->d_chunks, dim, c->d_bricks);
    c->launches += 2;
    // the host keeps the world too (VoxelBrickmap.get / is_walkable read the staging; edits go through it)
    UVT_CUDA(c, cudaMemcpyAsync(c->h_chunks, c->d_chunks, n_chunks * 4, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaMemcpyAsync(c->h_bricks, c->d_bricks, n * 2048, cudaMemcpyDeviceToHost, c->stream));
    UVT_CUDA(c, cudaStreamSynchronize(c->stream));
    UVT_CUDA(c, cudaGetLastError());
    procgen_release(c);
    return UVT_OK;
}

}  // extern "C"
