// MagicaVoxel .vox reader — stands in for zvox.VoxFile.from_reader, which the reference
// uses at src/engine/voxel.zig:119 and which is NOT vendored (zvox @deb43ad6, build.zig.zon:27-30).
//
// What the reference consumes (voxel.zig:94-127): `.models[]{size, voxels[]{x,y,z,color}}`
// in file order and `.palette.colors[256]` indexed with `color - 1`.  The container is the
// published MagicaVoxel RIFF-like format: "VOX " + version + MAIN whose
// children are SIZE/XYZI pairs, scene-graph chunks (skipped), RGBA (256 x 4 bytes) and
// optional IMAP/MATL/rOBJ/rCAM/NOTE (skipped; IMAP is NOT applied, the reference indexes
// the raw palette).  PARITY UNPINNED for zvox itself: no copy of it exists here.
#include "uvt_host.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

struct uvt_vox_model {
    uint32_t size[3];
    std::vector<uvt_vox_voxel> voxels;
};

struct uvt_vox_file {
    uint32_t version = 0;
    std::vector<uvt_vox_model> models;
    uint32_t palette[256];
    bool has_palette = false;
};

namespace {

thread_local std::string g_vox_error;

int fail(const char *msg) {
    g_vox_error = msg;
    return UVT_ERR_FORMAT;
}

inline uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

}  // namespace

extern "C" {

const char *uvt_vox_error(void) { return g_vox_error.c_str(); }

int uvt_vox_parse(const void *bytes, size_t n, uvt_vox_file **out) {
    if (!bytes || !out) return UVT_ERR_INVALID;
    const uint8_t *b = (const uint8_t *)bytes;
    if (n < 20 || std::memcmp(b, "VOX ", 4) != 0) return fail("not a .vox file (missing 'VOX ' magic)");
    uvt_vox_file *f = new uvt_vox_file;
    f->version = rd32(b + 4);
    std::memset(f->palette, 0, sizeof f->palette);

    if (std::memcmp(b + 8, "MAIN", 4) != 0) { delete f; return fail("first chunk is not MAIN"); }
    const uint64_t main_content = rd32(b + 12), main_children = rd32(b + 16);
    uint64_t off = 20 + main_content;
    const uint64_t end = off + main_children;
    if (end > n) { delete f; return fail("MAIN chunk overruns the file"); }

    bool have_size = false;
    uint32_t pending_size[3] = {0, 0, 0};
    while (off + 12 <= end) {
        const uint8_t *hdr = b + off;
        const uint64_t content = rd32(hdr + 4), children = rd32(hdr + 8);
        const uint8_t *body = hdr + 12;
        if (off + 12 + content + children > end) { delete f; return fail("chunk overruns MAIN"); }
        if (std::memcmp(hdr, "SIZE", 4) == 0) {
            if (content < 12) { delete f; return fail("short SIZE chunk"); }
            pending_size[0] = rd32(body);
            pending_size[1] = rd32(body + 4);
            pending_size[2] = rd32(body + 8);
            have_size = true;
        } else if (std::memcmp(hdr, "XYZI", 4) == 0) {
            if (!have_size) { delete f; return fail("XYZI without a preceding SIZE"); }
            if (content < 4) { delete f; return fail("short XYZI chunk"); }
            const uint64_t count = rd32(body);
            if (4 + count * 4 > content) { delete f; return fail("XYZI voxel count overruns its chunk"); }
            uvt_vox_model m;
            std::memcpy(m.size, pending_size, sizeof m.size);
            m.voxels.resize(count);
            for (uint64_t i = 0; i < count; ++i) {
                const uint8_t *v = body + 4 + i * 4;
                m.voxels[i] = uvt_vox_voxel{v[0], v[1], v[2], v[3]};
            }
            f->models.push_back(std::move(m));
            have_size = false;
        } else if (std::memcmp(hdr, "RGBA", 4) == 0) {
            if (content < 1024) { delete f; return fail("short RGBA chunk"); }
            for (int i = 0; i < 256; ++i) f->palette[i] = rd32(body + 4 * i);  // R | G<<8 | B<<16 | A<<24
            f->has_palette = true;
        }
        // everything else (PACK, nTRN, nGRP, nSHP, LAYR, IMAP, MATL, rOBJ, rCAM, NOTE, ...) is skipped
        off += 12 + content + children;
    }
    if (f->models.empty()) { delete f; return fail("no SIZE/XYZI model in file"); }
    if (!f->has_palette) { delete f; return fail("no RGBA chunk (MagicaVoxel default palette is not built in)"); }
    *out = f;
    return UVT_OK;
}

int uvt_vox_open(const char *path, uvt_vox_file **out) {
    FILE *fp = std::fopen(path, "rb");
    if (!fp) { g_vox_error = std::string("cannot open ") + (path ? path : "(null)"); return UVT_ERR_IO; }
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t got;
    while ((got = std::fread(tmp, 1, sizeof tmp, fp)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    std::fclose(fp);
    return uvt_vox_parse(buf.data(), buf.size(), out);
}

void uvt_vox_free(uvt_vox_file *f) { delete f; }
uint32_t uvt_vox_n_models(const uvt_vox_file *f) { return (uint32_t)f->models.size(); }
int uvt_vox_model_size(const uvt_vox_file *f, uint32_t model, uint32_t size_xyz[3]) {
    if (model >= f->models.size()) return UVT_ERR_INVALID;
    std::memcpy(size_xyz, f->models[model].size, sizeof(uint32_t) * 3);
    return UVT_OK;
}
uint32_t uvt_vox_model_n_voxels(const uvt_vox_file *f, uint32_t model) {
    return model < f->models.size() ? (uint32_t)f->models[model].voxels.size() : 0;
}
const uvt_vox_voxel *uvt_vox_model_voxels(const uvt_vox_file *f, uint32_t model) {
    return model < f->models.size() ? f->models[model].voxels.data() : nullptr;
}
const uint32_t *uvt_vox_palette(const uvt_vox_file *f) { return f->palette; }

}  // extern "C"
