// VoxelModelAtlas, host side: src/engine/voxel.zig:84-132.
//
// A 256^3 RGBA8 3-D atlas of 8^3 block models; model i lives in the 8^3 box at
// ((i%32)*8, ((i/32)%32)*8, (i/1024)*8) (voxel.zig:100-102) and is read by the shader at
// ((m&31)*8, ((m/32)&31)*8, ((m/1024)&31)*8) (assets/shaders/map.glsl:57-60).
// The atlas keeps a host copy of each loaded model and, when attached to a ctx, forwards
// the 8^3 sub-box through uvt_atlas_upload exactly as Texture.set_data_offset would
// (src/engine/graphics/texture.zig:70-72).
#include "uvt_host.h"

#include <array>
#include <cstring>
#include <vector>

struct uvt_atlas {
    uvt_ctx *ctx = nullptr;
    uvt_group *group = nullptr;
    size_t current_index = 0;  // voxel.zig:86
    std::vector<std::array<uint32_t, 512>> models;
};

namespace {

int push_model(uvt_atlas *a, const uint32_t texels[512]) {
    const size_t idx = a->current_index;
    const size_t base_x = idx % 32, base_y = (idx / 32) % 32, base_z = (idx / 1024) % 1024;
    if (base_z >= 32) return UVT_ERR_INVALID;  // outside the 256^3 texture (a GL error in the reference)
    if (a->ctx || a->group) {
        int rc = a->group ? uvt_group_atlas_upload(a->group, (uint32_t)base_x * 8, (uint32_t)base_y * 8, (uint32_t)base_z * 8, 8, 8, 8, texels)
                          : uvt_atlas_upload(a->ctx, (uint32_t)base_x * 8, (uint32_t)base_y * 8, (uint32_t)base_z * 8, 8, 8, 8, texels);
        if (rc != UVT_OK) return rc;
    }
    std::array<uint32_t, 512> m;
    std::memcpy(m.data(), texels, sizeof(uint32_t) * 512);
    a->models.push_back(m);
    a->current_index += 1;
    return UVT_OK;
}

// load_single_block_model: voxel.zig:94-112
int load_single(uvt_atlas *a, const uvt_vox_file *f, uint32_t mi) {
    uint32_t storage[512];
    std::memset(storage, 0, sizeof storage);
    const uint32_t *palette = uvt_vox_palette(f);
    const uvt_vox_voxel *vox = uvt_vox_model_voxels(f, mi);
    const uint32_t n = uvt_vox_model_n_voxels(f, mi);
    for (uint32_t i = 0; i < n; ++i) {
        const uvt_vox_voxel v = vox[i];
        if (v.x >= 8 || v.y >= 8 || v.z >= 8) return UVT_ERR_FORMAT;  // "assumed to be 8x8x8" (voxel.zig:114)
        if (v.color == 0) return UVT_ERR_FORMAT;                      // colors[color - 1] would underflow
        // MagicaVoxel is z-up: texel (x, y, z) <- voxel (x, z_mv, y_mv)
        storage[v.x + 8 * (v.z + 8 * v.y)] = palette[v.color - 1];
    }
    return push_model(a, storage);
}

int load_all(uvt_atlas *a, uvt_vox_file *f) {
    int rc = UVT_OK;
    const uint32_t n = uvt_vox_n_models(f);
    for (uint32_t i = 0; i < n && rc == UVT_OK; ++i) rc = load_single(a, f, i);
    uvt_vox_free(f);
    return rc;
}

}  // namespace

extern "C" {

int uvt_atlas_create(uvt_ctx *ctx, uvt_atlas **out) {
    if (!out) return UVT_ERR_INVALID;
    uvt_atlas *a = new uvt_atlas;
    a->ctx = ctx;
    *out = a;
    return UVT_OK;
}

int uvt_atlas_create_group(uvt_group *group, uvt_atlas **out) {
    if (!out || !group) return UVT_ERR_INVALID;
    uvt_atlas *a = new uvt_atlas;
    a->group = group;
    *out = a;
    return UVT_OK;
}

void uvt_atlas_destroy(uvt_atlas *a) { delete a; }

int uvt_atlas_load_block_model(uvt_atlas *a, const char *path) {
    uvt_vox_file *f = nullptr;
    int rc = uvt_vox_open(path, &f);
    if (rc != UVT_OK) return rc;
    return load_all(a, f);
}

int uvt_atlas_load_block_model_mem(uvt_atlas *a, const void *bytes, size_t n) {
    uvt_vox_file *f = nullptr;
    int rc = uvt_vox_parse(bytes, n, &f);
    if (rc != UVT_OK) return rc;
    return load_all(a, f);
}

int uvt_atlas_append_model(uvt_atlas *a, const uint32_t texels[512]) { return push_model(a, texels); }

uint32_t uvt_atlas_current_index(const uvt_atlas *a) { return (uint32_t)a->current_index; }

int uvt_atlas_get_model(const uvt_atlas *a, uint32_t idx, uint32_t texels[512]) {
    if (idx >= a->models.size()) return UVT_ERR_INVALID;
    std::memcpy(texels, a->models[idx].data(), sizeof(uint32_t) * 512);
    return UVT_OK;
}

}  // extern "C"
