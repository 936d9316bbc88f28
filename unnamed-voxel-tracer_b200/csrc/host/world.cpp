// VoxelBrickmap + GpuBlockAllocator + procgen, host side.
//
// Mirrors src/engine/voxel.zig:25-82 (VoxelBrickmap), src/engine/graphics/
// gpu_block_allocator.zig:4-41 (bump allocator with x2 growth) and src/procgen.zig:6-70.
// When attached to a uvt_ctx the storage is the ctx's pinned staging (uvt_world_alloc /
// uvt_world_grow), the analogue of the reference's persistently mapped GL buffers.
#include "uvt_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <new>
#include <vector>

struct uvt_brickmap {
    uvt_ctx *ctx = nullptr;
    uvt_group *group = nullptr;  // attached to a multi-GPU group instead of one ctx (ctx = its member 0)
    uint32_t dim = 0;       // blocks per axis
    uint32_t cdim = 0;      // chunks per axis = dim / 8
    uint32_t *chunks = nullptr;
    uint32_t *bricks = nullptr;
    size_t block_index = 0;      // next free brick
    size_t max_block_index = 0;  // capacity in bricks
    // blocks written since the last bind (uvt_brickmap_bind publishes only what changed)
    bool bound = false, dirty = true, dirty_all = true;
    uint32_t dirty_lo[3] = {0, 0, 0}, dirty_hi[3] = {0, 0, 0};
};

namespace {

constexpr uint32_t kChunk = 8;
constexpr size_t kBrickWords = 512;

inline size_t pos_to_index(size_t dim, size_t x, size_t y, size_t z) { return x + dim * (y + z * dim); }

// GpuBlockAllocator.alloc: gpu_block_allocator.zig:20-30
int brick_alloc(uvt_brickmap *bm, size_t *out) {
    if (bm->block_index >= bm->max_block_index) {
        size_t new_cap = bm->max_block_index * 2;
        if (bm->ctx) {
            uint32_t *nb = nullptr;
            int rc = bm->group ? uvt_group_world_grow(bm->group, new_cap, &nb) : uvt_world_grow(bm->ctx, new_cap, &nb);
            if (rc != UVT_OK) return rc;
            bm->bricks = nb;
        } else {
            uint32_t *nb = (uint32_t *)std::realloc(bm->bricks, new_cap * kBrickWords * sizeof(uint32_t));
            if (!nb) return UVT_ERR_OOM;
            // fresh blocks read as zero (the reference relies on GL zero-initialised storage)
            std::memset(nb + bm->max_block_index * kBrickWords, 0,
                        (new_cap - bm->max_block_index) * kBrickWords * sizeof(uint32_t));
            bm->bricks = nb;
        }
        bm->max_block_index = new_cap;
    }
    *out = bm->block_index++;
    return UVT_OK;
}

// VoxelBrickmap.get_block_for_chunk: voxel.zig:47-56
int block_for_chunk(uvt_brickmap *bm, size_t chx, size_t chy, size_t chz, size_t *out) {
    uint32_t &slot = bm->chunks[pos_to_index(bm->cdim, chx, chy, chz)];
    if (slot > 0) {
        *out = slot - 1;
        return UVT_OK;
    }
    size_t idx;
    int rc = brick_alloc(bm, &idx);
    if (rc != UVT_OK) return rc;
    // re-resolve: growth never moves `chunks`, only the brick pool
    bm->chunks[pos_to_index(bm->cdim, chx, chy, chz)] = (uint32_t)idx + 1;
    *out = idx;
    return UVT_OK;
}

}  // namespace

extern "C" {

static int brickmap_create(uvt_ctx *ctx, uvt_group *group, uint32_t dim, uvt_brickmap **out);

int uvt_brickmap_create(uvt_ctx *ctx, uint32_t dim, uvt_brickmap **out) { return brickmap_create(ctx, nullptr, dim, out); }

int uvt_brickmap_create_group(uvt_group *group, uint32_t dim, uvt_brickmap **out) {
    if (!group) return UVT_ERR_INVALID;
    return brickmap_create(uvt_group_member(group, 0), group, dim, out);
}

static int brickmap_create(uvt_ctx *ctx, uvt_group *group, uint32_t dim, uvt_brickmap **out) {
    if (!out || dim == 0 || dim % kChunk != 0 || dim > 4096) return UVT_ERR_INVALID;  // 4096: the bound of uvt_world_alloc
    uvt_brickmap *bm = new (std::nothrow) uvt_brickmap;
    if (!bm) return UVT_ERR_OOM;
    bm->ctx = ctx;
    bm->group = group;
    bm->dim = dim;
    bm->cdim = dim / kChunk;
    const size_t n_chunks = (size_t)bm->cdim * bm->cdim * bm->cdim;
    const size_t cap = dim;  // GpuBlockAllocator.init(dim): voxel.zig:36
    if (ctx) {
        int rc = group ? uvt_group_world_alloc(group, dim, &bm->chunks, &bm->bricks, cap) : uvt_world_alloc(ctx, dim, &bm->chunks, &bm->bricks, cap);
        if (rc != UVT_OK) { delete bm; return rc; }
    } else {
        bm->chunks = (uint32_t *)std::calloc(n_chunks, sizeof(uint32_t));
        bm->bricks = (uint32_t *)std::calloc(cap * kBrickWords, sizeof(uint32_t));
        if (!bm->chunks || !bm->bricks) { std::free(bm->chunks); std::free(bm->bricks); delete bm; return UVT_ERR_OOM; }
    }
    bm->max_block_index = cap;
    *out = bm;
    return UVT_OK;
}

void uvt_brickmap_destroy(uvt_brickmap *bm) {
    if (!bm) return;
    if (!bm->ctx) { std::free(bm->chunks); std::free(bm->bricks); }  // ctx staging is owned by the ctx
    delete bm;
}

void uvt_brickmap_clear(uvt_brickmap *bm) {
    const size_t n_chunks = (size_t)bm->cdim * bm->cdim * bm->cdim;
    std::memset(bm->chunks, 0, n_chunks * sizeof(uint32_t));
    bm->block_index = 0;
    std::memset(bm->bricks, 0, bm->max_block_index * kBrickWords * sizeof(uint32_t));
    bm->dirty = bm->dirty_all = true;
}

void uvt_brickmap_mark_dirty(uvt_brickmap *bm) {
    if (bm) bm->dirty = bm->dirty_all = true;
}

int uvt_brickmap_set(uvt_brickmap *bm, uint32_t x, uint32_t y, uint32_t z, uint32_t voxel) {
    if (x >= bm->dim || y >= bm->dim || z >= bm->dim) return UVT_ERR_INVALID;  // the reference would index out of bounds
    size_t blk;
    int rc = block_for_chunk(bm, x / kChunk, y / kChunk, z / kChunk, &blk);
    if (rc != UVT_OK) return rc;
    bm->bricks[blk * kBrickWords + (x % kChunk) + ((y % kChunk) + (z % kChunk) * kChunk) * kChunk] = voxel;
    const uint32_t p[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        bm->dirty_lo[a] = bm->dirty ? std::min(bm->dirty_lo[a], p[a]) : p[a];
        bm->dirty_hi[a] = bm->dirty ? std::max(bm->dirty_hi[a], p[a]) : p[a];
    }
    bm->dirty = true;
    return UVT_OK;
}

uint32_t uvt_brickmap_get(uvt_brickmap *bm, uint32_t x, uint32_t y, uint32_t z) {
    if (x >= bm->dim || y >= bm->dim || z >= bm->dim) return 0;
    uint32_t index = bm->chunks[pos_to_index(bm->cdim, x / kChunk, y / kChunk, z / kChunk)];
    if (index == 0) return 0;
    return bm->bricks[(size_t)(index - 1) * kBrickWords + (x % kChunk) + ((y % kChunk) + (z % kChunk) * kChunk) * kChunk];
}

int uvt_brickmap_is_walkable(uvt_brickmap *bm, uint32_t x, uint32_t y, uint32_t z) {
    uint32_t v = uvt_brickmap_get(bm, x, y, z);
    return ((v & UVT_VOXEL_SOLID) == 0) || v == 0;
}

uint32_t uvt_brickmap_dim(const uvt_brickmap *bm) { return bm->dim; }
size_t uvt_brickmap_n_bricks(const uvt_brickmap *bm) { return bm->block_index; }
size_t uvt_brickmap_capacity(const uvt_brickmap *bm) { return bm->max_block_index; }
const uint32_t *uvt_brickmap_chunks(const uvt_brickmap *bm) { return bm->chunks; }
const uint32_t *uvt_brickmap_bricks(const uvt_brickmap *bm) { return bm->bricks; }

int uvt_brickmap_bind(uvt_brickmap *bm) {
    if (!bm->ctx) return UVT_ERR_INVALID;
    // the reference's mapping is live and bind() is a per-frame GL call (game.zig:236): publish only what was written
    int rc = UVT_OK;
    if (!bm->bound || bm->dirty_all)
        rc = bm->group ? uvt_group_world_commit(bm->group, bm->block_index) : uvt_world_commit(bm->ctx, bm->block_index);
    else if (bm->dirty)
        rc = bm->group ? uvt_group_world_commit_region(bm->group, bm->block_index, bm->dirty_lo, bm->dirty_hi)
                       : uvt_world_commit_region(bm->ctx, bm->block_index, bm->dirty_lo, bm->dirty_hi);
    if (rc != UVT_OK) return rc;
    bm->bound = true;
    bm->dirty = bm->dirty_all = false;
    return UVT_OK;
}

int uvt_brickmap_save(const uvt_brickmap *bm, const char *path) {
    FILE *f = std::fopen(path, "wb");
    if (!f) return UVT_ERR_IO;
    const uint32_t hdr[4] = {0x57545655u /* "UVTW" */, 1u, bm->dim, (uint32_t)bm->block_index};
    const size_t n_chunks = (size_t)bm->cdim * bm->cdim * bm->cdim;
    bool ok = std::fwrite(hdr, sizeof hdr, 1, f) == 1 &&
              std::fwrite(bm->chunks, sizeof(uint32_t), n_chunks, f) == n_chunks &&
              std::fwrite(bm->bricks, sizeof(uint32_t) * kBrickWords, bm->block_index, f) == bm->block_index;
    std::fclose(f);
    return ok ? UVT_OK : UVT_ERR_IO;
}

int uvt_brickmap_load(uvt_ctx *ctx, const char *path, uvt_brickmap **out) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return UVT_ERR_IO;
    uint32_t hdr[4];
    if (std::fread(hdr, sizeof hdr, 1, f) != 1 || hdr[0] != 0x57545655u || hdr[1] != 1u) { std::fclose(f); return UVT_ERR_FORMAT; }
    if (hdr[2] == 0 || hdr[2] % kChunk != 0 || hdr[2] > 4096) { std::fclose(f); return UVT_ERR_FORMAT; }
    {   // the header must agree with the file: nothing is allocated for a size the file cannot back
        const size_t cdim = hdr[2] / kChunk, want = sizeof hdr + cdim * cdim * cdim * sizeof(uint32_t) + (size_t)hdr[3] * kBrickWords * sizeof(uint32_t);
        if (std::fseek(f, 0, SEEK_END) != 0) { std::fclose(f); return UVT_ERR_IO; }
        const long size = std::ftell(f);
        if (size < 0 || (size_t)size < want || (size_t)hdr[3] > cdim * cdim * cdim || std::fseek(f, (long)sizeof hdr, SEEK_SET) != 0) {
            std::fclose(f);
            return UVT_ERR_FORMAT;
        }
    }
    uvt_brickmap *bm = nullptr;
    int rc = uvt_brickmap_create(ctx, hdr[2], &bm);
    if (rc != UVT_OK) { std::fclose(f); return rc; }
    const size_t n_chunks = (size_t)bm->cdim * bm->cdim * bm->cdim;
    const size_t n_bricks = hdr[3];
    while (bm->max_block_index < n_bricks) {  // grow exactly like the allocator would
        bm->block_index = bm->max_block_index;
        size_t dummy;
        rc = brick_alloc(bm, &dummy);
        if (rc != UVT_OK) { std::fclose(f); uvt_brickmap_destroy(bm); return rc; }
    }
    bool ok = std::fread(bm->chunks, sizeof(uint32_t), n_chunks, f) == n_chunks &&
              std::fread(bm->bricks, sizeof(uint32_t) * kBrickWords, n_bricks, f) == n_bricks;
    std::fclose(f);
    if (!ok) { uvt_brickmap_destroy(bm); return UVT_ERR_FORMAT; }
    {   // a chunk entry is 0 or brick + 1 of a brick the file holds, and no two chunks share a brick:
        // get/set index the pool through these entries without a further check
        std::vector<uint8_t> seen(n_bricks, 0);
        for (size_t i = 0; i < n_chunks; ++i) {
            const uint32_t e = bm->chunks[i];
            if (e == 0) continue;
            if (e > n_bricks || seen[e - 1]) { uvt_brickmap_destroy(bm); return UVT_ERR_FORMAT; }
            seen[e - 1] = 1;
        }
    }
    bm->block_index = n_bricks;
    *out = bm;
    return UVT_OK;
}

// ---- procgen: src/procgen.zig:6-70 -----------------------------------------------------

// place_tree: procgen.zig:55-70
static int place_tree(uvt_lcg *lcg, uvt_brickmap *w, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t trunk_height = uvt_lcg_rand(lcg) % 4 + 4;
    int rc = uvt_brickmap_set(w, x + 1, y, z + 1, uvt_voxel(15, 1));
    for (uint32_t off = 0; off < trunk_height && rc == UVT_OK; ++off)
        rc = uvt_brickmap_set(w, x + 1, y + off, z + 1, uvt_voxel(14 + uvt_lcg_rand(lcg) % 3, 1));
    for (uint32_t a = 0; a < 3; ++a)
        for (uint32_t b = 0; b < 3; ++b)
            for (uint32_t c = 0; c < 3 && rc == UVT_OK; ++c)
                rc = uvt_brickmap_set(w, x + a, y + trunk_height + b, z + c, uvt_voxel(18 + uvt_lcg_rand(lcg) % 2, 1));
    return rc;
}

int uvt_procgen(uvt_brickmap *world, uint32_t dim, float offset_x, float offset_y) {
    if (!world || dim != world->dim) return UVT_ERR_INVALID;
    uvt_lcg lcg{0x46AE4F};
    int rc = UVT_OK;

    // water slab (procgen.zig:10-19)
    for (uint32_t x = 0; x < dim; ++x)
        for (uint32_t z = 0; z < dim; ++z)
            for (uint32_t y = 0; y < 16; ++y)
                if ((rc = uvt_brickmap_set(world, x, y, z, uvt_voxel(13, 1))) != UVT_OK) return rc;

    for (uint32_t x = 0; x < dim; ++x) {
        for (uint32_t z = 0; z < dim; ++z) {
            const uint32_t vh = uvt_procgen_height(dim, x, z, offset_x, offset_y);

            for (uint32_t h = 0; h < vh; ++h) {
                if ((rc = uvt_brickmap_set(world, x, h, z, uvt_voxel(21 + uvt_lcg_rand(&lcg) % 3, 1))) != UVT_OK) return rc;
                if (h <= 15) {
                    rc = uvt_brickmap_set(world, x, h, z, uvt_voxel(25 + uvt_lcg_rand(&lcg) % 3, 1));
                } else if (h == vh - 1 && h > 15) {
                    rc = uvt_brickmap_set(world, x, h, z, uvt_voxel(uvt_lcg_rand(&lcg) % 6, 1));
                }
                if (rc != UVT_OK) return rc;
            }

            if (vh > 16) {
                // `continue` at procgen.zig:37-38 also skips the trailing draw
                if (uvt_brickmap_get(world, x, vh, z) != 0) continue;

                if (uvt_lcg_rand(&lcg) % 5 == 0)
                    if ((rc = uvt_brickmap_set(world, x, vh, z, uvt_voxel(7 + uvt_lcg_rand(&lcg) % 5, 0))) != UVT_OK) return rc;

                if (uvt_lcg_rand(&lcg) % 71 == 0)
                    if ((rc = uvt_brickmap_set(world, x, vh, z, uvt_voxel(7 + 5, 1))) != UVT_OK) return rc;

                // the reference's literal 500-block guard (procgen.zig:47); maps narrower than 503 blocks additionally keep the
                // 3x3 canopy (x..x+2, z..z+2) inside the map, where the reference would index out of bounds (same at dim >= 503)
                if (uvt_lcg_rand(&lcg) % 420 == 0 && x < 500 && z < 500 && x > 5 && z > 5 && x + 2 < dim && z + 2 < dim)
                    if ((rc = place_tree(&lcg, world, x, vh, z)) != UVT_OK) return rc;
            }
            (void)uvt_lcg_rand(&lcg);
        }
    }
    return UVT_OK;
}

// procgen on the device: the same world as uvt_procgen above, byte for byte, generated by kernels
// (csrc/procgen.cuh).  The map must be attached to a ctx and empty.  The brick pool is grown exactly as
// GpuBlockAllocator.alloc would have grown it (doubling from `dim` bricks), so capacity and dump are identical too.
int uvt_procgen_device(uvt_brickmap *world, uint32_t dim, float offset_x, float offset_y) {
    if (!world || dim != world->dim || !world->ctx || world->group || world->block_index != 0) return UVT_ERR_INVALID;
    size_t n = 0;
    int rc = uvt_world_procgen_plan(world->ctx, offset_x, offset_y, &n);
    if (rc != UVT_OK) return rc;
    size_t new_cap = world->max_block_index;
    while (new_cap < n) new_cap *= 2;   // the capacity the doubling allocator ends at, reached in one step (the map is empty)
    if (new_cap != world->max_block_index) {
        uint32_t *nb = nullptr;
        rc = uvt_world_grow(world->ctx, new_cap, &nb);
        if (rc != UVT_OK) return rc;
        world->bricks = nb;
        world->max_block_index = new_cap;
    }
    rc = uvt_world_procgen_fill(world->ctx);
    if (rc != UVT_OK) return rc;
    world->block_index = n;
    world->dirty = world->dirty_all = true;
    return UVT_OK;
}

}  // extern "C"
