// uvt_group: several GPUs behind one handle, in ONE process.
//
// The reference is a single-process game (src/game.zig): its renderer cannot be launched one rank per GPU.
// A group owns one uvt_ctx per device.  The world and the atlas are replicated (one pinned host staging, committed to
// every member), the frame is cut into interleaved 16-row bands (member i renders bands i, i+n, ...) and
// every member's kernels store their finished bands straight into the frame of member 0 through peer access — the
// same fused compute + "gather" the torchrun path uses with CUDA IPC, without a second process.
//
// Built on the public C ABI only (include/uvt.h) plus the CUDA runtime for peer access and the shared frame.
#include "uvt.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <new>
#include <string>
#include <vector>

struct uvt_group {
    std::vector<uvt_ctx *> members;
    std::vector<int> devices;
    std::string error;
    uint32_t W = 0, H = 0;
    void *frame = nullptr;  // full W x H RGBA8 frame on devices[0]
};

namespace {

constexpr uint32_t kBandRows = 16;  // two CTA tile rows; 2160 rows over 8 members: 272 vs 256-288 rows with 32-row bands
std::string g_group_create_error;

int group_error(uvt_group *g, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    (g ? g->error : g_group_create_error) = buf;
    return code;
}

// status of member i failed: carry its message
int member_error(uvt_group *g, int i, int rc) {
    const char *msg = uvt_last_error(g->members[(size_t)i]);
    return group_error(g, rc, "member %d (device %d): %s", i, g->devices[(size_t)i], msg ? msg : "");
}

#define UVT_EACH(g, call)                                   \
    do {                                                    \
        for (int i_ = 0; i_ < (int)(g)->members.size(); ++i_) { \
            uvt_ctx *m = (g)->members[(size_t)i_];          \
            int rc_ = (call);                               \
            if (rc_ != UVT_OK) return member_error((g), i_, rc_); \
        }                                                   \
    } while (0)

void free_frame(uvt_group *g) {
    if (g->frame) {
        cudaSetDevice(g->devices[0]);
        cudaFree(g->frame);
        g->frame = nullptr;
    }
}

}  // namespace

extern "C" {

int uvt_group_create(const uvt_params *params, const int *devices, int n, uvt_group **out) {
    if (!out) return group_error(nullptr, UVT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n < 1 || n > 64) return group_error(nullptr, UVT_ERR_INVALID, "need 1..64 devices");
    uvt_group *g = new (std::nothrow) uvt_group;
    if (!g) return group_error(nullptr, UVT_ERR_OOM, "out of host memory");
    for (int i = 0; i < n; ++i) {
        uvt_ctx *c = nullptr;
        int rc = uvt_create(params, devices[i], &c);
        if (rc != UVT_OK) {
            const char *msg = uvt_last_error(nullptr);
            group_error(nullptr, rc, "member %d (device %d): %s", i, devices[i], msg ? msg : "");
            uvt_group_destroy(g);
            return rc;
        }
        g->members.push_back(c);
        g->devices.push_back(devices[i]);
    }
    // every member stores into the frame of member 0: map device 0's memory into the others
    for (int i = 1; i < n; ++i) {
        if (devices[i] == devices[0]) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[i], devices[0]);
        if (!can) {
            group_error(nullptr, UVT_ERR_CUDA, "device %d cannot access device %d's memory (no peer access)", devices[i], devices[0]);
            uvt_group_destroy(g);
            return UVT_ERR_CUDA;
        }
        cudaSetDevice(devices[i]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            group_error(nullptr, UVT_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[i], devices[0], cudaGetErrorString(e));
            uvt_group_destroy(g);
            return UVT_ERR_CUDA;
        }
        (void)cudaGetLastError();
    }
    *out = g;
    return UVT_OK;
}

void uvt_group_destroy(uvt_group *g) {
    if (!g) return;
    for (uvt_ctx *c : g->members) uvt_sync(c);
    free_frame(g);
    // the world staging belongs to member 0: detach the borrowers first
    for (size_t i = g->members.size(); i-- > 0;) uvt_destroy(g->members[i]);
    delete g;
}

int uvt_group_size(const uvt_group *g) { return g ? (int)g->members.size() : 0; }

uvt_ctx *uvt_group_member(uvt_group *g, int i) { return (g && i >= 0 && i < (int)g->members.size()) ? g->members[(size_t)i] : nullptr; }

const char *uvt_group_last_error(uvt_group *g) { return g ? g->error.c_str() : g_group_create_error.c_str(); }

// ---- world + atlas: one staging, every member commits from it ---------------------------------
int uvt_group_world_alloc(uvt_group *g, uint32_t dim, uint32_t **chunks_host, uint32_t **bricks_host, size_t brick_capacity) {
    if (!g) return UVT_ERR_INVALID;
    int rc = uvt_world_alloc(g->members[0], dim, chunks_host, bricks_host, brick_capacity);
    if (rc != UVT_OK) return member_error(g, 0, rc);
    for (int i = 1; i < (int)g->members.size(); ++i) {
        rc = uvt_world_use_staging(g->members[(size_t)i], dim, *chunks_host, *bricks_host, brick_capacity);
        if (rc != UVT_OK) return member_error(g, i, rc);
    }
    return UVT_OK;
}

int uvt_group_world_grow(uvt_group *g, size_t new_capacity, uint32_t **bricks_host) {
    if (!g) return UVT_ERR_INVALID;
    for (uvt_ctx *c : g->members) uvt_sync(c);  // nobody may still be reading the old staging
    int rc = uvt_world_grow(g->members[0], new_capacity, bricks_host);
    if (rc != UVT_OK) return member_error(g, 0, rc);
    for (int i = 1; i < (int)g->members.size(); ++i) {
        rc = uvt_world_use_staging(g->members[(size_t)i], 0, nullptr, *bricks_host, new_capacity);  // dim 0: keep the chunk table, swap the pool
        if (rc != UVT_OK) return member_error(g, i, rc);
    }
    return UVT_OK;
}

int uvt_group_world_commit(uvt_group *g, size_t n_bricks) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_world_commit(m, n_bricks));
    return UVT_OK;
}

int uvt_group_world_commit_region(uvt_group *g, size_t n_bricks, const uint32_t lo[3], const uint32_t hi[3]) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_world_commit_region(m, n_bricks, lo, hi));
    return UVT_OK;
}

int uvt_group_atlas_upload(uvt_group *g, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t w, uint32_t h, uint32_t d, const uint32_t *rgba) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_atlas_upload(m, ox, oy, oz, w, h, d, rgba));
    return UVT_OK;
}

int uvt_group_set_entity_mode(uvt_group *g, uint32_t mode) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_set_entity_mode(m, mode));
    return UVT_OK;
}

int uvt_group_set_entities(uvt_group *g, const float *positions_xyz, uint32_t n) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_set_entities(m, positions_xyz, n));
    return UVT_OK;
}

int uvt_group_entity_model_upload(uvt_group *g, uint32_t size, const uint32_t *rgba, uint32_t max_steps) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_entity_model_upload(m, size, rgba, max_steps));
    return UVT_OK;
}

// ---- per frame ------------------------------------------------------------------------------------
int uvt_group_set_camera(uvt_group *g, const uvt_camera *cam) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_set_camera(m, cam));
    return UVT_OK;
}

int uvt_group_resize(uvt_group *g, uint32_t width, uint32_t height) {
    if (!g) return UVT_ERR_INVALID;
    const int n = (int)g->members.size();
    for (uvt_ctx *c : g->members) uvt_sync(c);
    free_frame(g);
    g->W = width;
    g->H = height;
    if (n == 1) {
        int rc = uvt_resize(g->members[0], width, height);
        return rc == UVT_OK ? UVT_OK : member_error(g, 0, rc);
    }
    if (width && height) {
        cudaSetDevice(g->devices[0]);
        cudaError_t e = cudaMalloc(&g->frame, (size_t)width * height * 4);
        if (e == cudaSuccess) e = cudaMemset(g->frame, 0, (size_t)width * height * 4);
        if (e != cudaSuccess) return group_error(g, e == cudaErrorMemoryAllocation ? UVT_ERR_OOM : UVT_ERR_CUDA, "shared frame: %s", cudaGetErrorString(e));
    }
    for (int i = 0; i < n; ++i) {
        uvt_ctx *m = g->members[(size_t)i];
        int rc = uvt_set_partition(m, kBandRows, (uint32_t)n, (uint32_t)i);
        if (rc == UVT_OK) rc = uvt_resize(m, width, height);
        if (rc == UVT_OK) rc = uvt_bind_frame_target(m, g->frame, 0, 1);
        if (rc != UVT_OK) return member_error(g, i, rc);
    }
    return UVT_OK;
}

int uvt_group_dispatch_frame(uvt_group *g) {
    if (!g) return UVT_ERR_INVALID;
    // pass by pass rather than member by member: every device starts after n launches instead of 2n (all asynchronous)
    UVT_EACH(g, uvt_dispatch_primary(m));
    UVT_EACH(g, uvt_dispatch_secondary_shade(m));  // shadow pass + blit (with the peer stores of the bands) in one launch per member
    return UVT_OK;
}

int uvt_group_sync(uvt_group *g) {
    if (!g) return UVT_ERR_INVALID;
    UVT_EACH(g, uvt_sync(m));
    return UVT_OK;
}

int uvt_group_readback_frame(uvt_group *g, void *dst, size_t bytes) {
    if (!g || !dst) return UVT_ERR_INVALID;
    if (bytes != (size_t)g->W * g->H * 4) return group_error(g, UVT_ERR_INVALID, "frame is %zu bytes, caller passed %zu", (size_t)g->W * g->H * 4, bytes);
    UVT_EACH(g, uvt_sync(m));
    int rc = g->members.size() == 1 ? uvt_readback(g->members[0], UVT_BUF_FRAME, dst, bytes) : uvt_read_device(g->members[0], g->frame, dst, bytes);
    return rc == UVT_OK ? UVT_OK : member_error(g, 0, rc);
}

int uvt_group_frame_ptr(uvt_group *g, void **dptr) {
    if (!g || !dptr) return UVT_ERR_INVALID;
    if (g->members.size() == 1) {
        int rc = uvt_device_ptr(g->members[0], UVT_BUF_FRAME, dptr);
        return rc == UVT_OK ? UVT_OK : member_error(g, 0, rc);
    }
    *dptr = g->frame;
    return g->frame ? UVT_OK : group_error(g, UVT_ERR_INVALID, "no frame (uvt_group_resize first)");
}

int uvt_group_count_pass(uvt_group *g, int which, uvt_counters *sum) {
    if (!g || !sum) return UVT_ERR_INVALID;
    uvt_counters total = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < (int)g->members.size(); ++i) {
        uvt_counters c;
        int rc = uvt_count_pass(g->members[(size_t)i], which, &c);
        if (rc != UVT_OK) return member_error(g, i, rc);
        total.rays += c.rays; total.t_in += c.t_in; total.t_chunk += c.t_chunk;
        total.t_block += c.t_block; total.hits += c.hits; total.early_out += c.early_out;
    }
    *sum = total;
    return UVT_OK;
}

}  // extern "C"
