// harness.cpp — runs the reference's own shader text (translated by translate.py, compiled against glsl_shim.h) over a
// frame on the CPU.  TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libglslref.so, loaded by oracle/glslref.py, used by
// tests/ to pin oracle/oracle.c (and through it the CUDA path) against the reference's source.
//
// The harness plays the part of the GL state: it binds the camera uniform block (binding 8), the brick pool (9), the
// chunk table (10), the 256^3 model atlas (image unit 6) and the G-buffer images (0-3) exactly as src/game.zig:235-252
// does, then invokes main() once per invocation of the reference's dispatch geometry ((W/32+1) x (H/32+1) groups of
// 32 x 32, src/game.zig:241-248) or once per fragment of the full-screen quad (blit; texPos = fragment centre / size,
// which is what blit.vertex.glsl:5-14 interpolates to).
#include "glsl_shim.h"

#include <cstddef>
#ifdef _OPENMP
#include <omp.h>
#endif

#define UVT_SHADER_NS(ns, file)                                  \
    namespace glsl { namespace ns {                              \
    static thread_local uvec3 gl_GlobalInvocationID(0u, 0u, 0u); \
    static int ref_map_dimension = 512;                          \
    }}                                                           \
    namespace glsl { namespace ns {
#define UVT_SHADER_END }}

UVT_SHADER_NS(primary, 0)
#ifdef UVT_GLSL_ENTITIES
#include "primary_entities.gen.inc"
#else
#include "primary.gen.inc"
#endif
UVT_SHADER_END
UVT_SHADER_NS(secondary, 0)
#ifdef UVT_GLSL_ENTITIES
#include "secondary_entities.gen.inc"
#else
#include "secondary.gen.inc"
#endif
UVT_SHADER_END
UVT_SHADER_NS(blit, 0)
#include "blit.gen.inc"
UVT_SHADER_END

using namespace glsl;

namespace {
struct Camera {  // camera.zig:12-16 == camera.glsl:2-6 (std140)
    float pos[4];
    float mat[16];
    float fov;
    float pad[3];
};

template <class F>
void for_each_invocation(uint32_t W, uint32_t H, F f) {
    const int64_t gx = W / 32 + 1, gy = H / 32 + 1;  // game.zig:241-242
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int64_t by = 0; by < gy; ++by)
        for (int64_t bx = 0; bx < gx; ++bx)
            for (uint32_t ly = 0; ly < 32; ++ly)
                for (uint32_t lx = 0; lx < 32; ++lx) f((uint32_t)bx * 32u + lx, (uint32_t)by * 32u + ly);
}
}  // namespace

extern "C" {

const char *ref_info(void) {
#ifdef UVT_GLSL_ENTITIES
    return "reference GLSL (primary.comp, secondary.comp, blit.fragment + camera/map/rng) with map.glsl:199 deleted and primary.comp.glsl:47-54 uncommented";
#else
    return "reference GLSL (primary.comp, secondary.comp, blit.fragment + camera/map/rng), text unmodified";
#endif
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// primary.comp.glsl main over the whole G-buffer
void ref_primary(uint32_t dim, const uint32_t *chunks, const uint32_t *bricks, const uint32_t *atlas256, const void *camera, uint32_t W, uint32_t H,
                 uint32_t *albedo, uint32_t *normal, float *position) {
    namespace P = glsl::primary;
    const Camera *cam = static_cast<const Camera *>(camera);
    P::ref_map_dimension = (int)dim;
    P::chunks = const_cast<uint32_t *>(chunks);
    P::data = const_cast<uint32_t *>(bricks);
    P::model = image3D{256, 256, 256, atlas256};
    P::C_position = vec4(cam->pos[0], cam->pos[1], cam->pos[2], cam->pos[3]);
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) P::C_view.c[j][i] = cam->mat[4 * j + i];  // 4 consecutive floats = one GLSL column
    P::fov = cam->fov;
    P::frameColor = image2D{(int)W, (int)H, FMT_RGBA8, albedo};
    P::frameNormal = image2D{(int)W, (int)H, FMT_RGBA8, normal};
    P::framePosition = image2D{(int)W, (int)H, FMT_RGBA32F, position};
    for_each_invocation(W, H, [](uint32_t x, uint32_t y) {
        P::gl_GlobalInvocationID = uvec3(x, y, 0u);
        P::shader_main();
    });
}

// secondary.comp.glsl main
void ref_secondary(uint32_t dim, const uint32_t *chunks, const uint32_t *bricks, const uint32_t *atlas256, uint32_t W, uint32_t H,
                   const uint32_t *albedo, const uint32_t *normal, const float *position, uint32_t *illum) {
    namespace S = glsl::secondary;
    S::ref_map_dimension = (int)dim;
    S::chunks = const_cast<uint32_t *>(chunks);
    S::data = const_cast<uint32_t *>(bricks);
    S::model = image3D{256, 256, 256, atlas256};
    S::frameColor = image2D{(int)W, (int)H, FMT_RGBA8, const_cast<uint32_t *>(albedo)};
    S::frameNormal = image2D{(int)W, (int)H, FMT_RGBA8, const_cast<uint32_t *>(normal)};
    S::framePosition = image2D{(int)W, (int)H, FMT_RGBA32F, const_cast<float *>(position)};
    S::frameIllumination = image2D{(int)W, (int)H, FMT_RGBA8, illum};
    for_each_invocation(W, H, [](uint32_t x, uint32_t y) {
        S::gl_GlobalInvocationID = uvec3(x, y, 0u);
        S::shader_main();
    });
}

// blit.fragment.glsl main for every fragment of the full-screen quad; fragColor goes to an RGBA8 target
void ref_blit(uint32_t W, uint32_t H, const uint32_t *albedo, const uint32_t *normal, const float *position, const uint32_t *illum, uint32_t *frame) {
    namespace B = glsl::blit;
    B::frameColor = image2D{(int)W, (int)H, FMT_RGBA8, const_cast<uint32_t *>(albedo)};
    B::frameNormal = image2D{(int)W, (int)H, FMT_RGBA8, const_cast<uint32_t *>(normal)};
    B::framePosition = image2D{(int)W, (int)H, FMT_RGBA32F, const_cast<float *>(position)};
    B::frameIllumination = image2D{(int)W, (int)H, FMT_RGBA8, const_cast<uint32_t *>(illum)};
    image2D target{(int)W, (int)H, FMT_RGBA8, frame};
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)H; ++y)
        for (uint32_t x = 0; x < W; ++x) {
            B::texPos = vec2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
            B::shader_main();
            imageStore(target, ivec2((int)x, (int)y), B::fragColor);
        }
}

// traceMap (map.glsl:83-168) for one ray: HitInfo as returned
void ref_trace_map(uint32_t dim, const uint32_t *chunks, const uint32_t *bricks, const uint32_t *atlas256, const float o[3], const float d[3],
                   int max_steps, uint32_t *data, float hit_pos[3], float normal[3]) {
    namespace P = glsl::primary;
    P::ref_map_dimension = (int)dim;
    P::chunks = const_cast<uint32_t *>(chunks);
    P::data = const_cast<uint32_t *>(bricks);
    P::model = image3D{256, 256, 256, atlas256};
    const P::HitInfo h = P::traceMap(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]), max_steps);
    *data = h.data;
    for (int k = 0; k < 3; ++k) { hit_pos[k] = h.hit_pos[k]; normal[k] = h.normal[k]; }
}

// traceEntities (map.glsl:172-248) for one ray
void ref_trace_entities(const uint32_t *atlas256, const float o[3], const float d[3], float max_distance, uint32_t *data, float hit_pos[3], float normal[3]) {
    namespace P = glsl::primary;
    P::model = image3D{256, 256, 256, atlas256};
    const P::HitInfo h = P::traceEntities(vec3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]), max_distance);
    *data = h.data;
    for (int k = 0; k < 3; ++k) { hit_pos[k] = h.hit_pos[k]; normal[k] = h.normal[k]; }
}

// SkyDome2 (camera.glsl:11-19)
void ref_sky_dome2(const float rd[3], float rgba[4]) {
    const vec4 c = glsl::primary::SkyDome2(vec3(0.0f), vec3(rd[0], rd[1], rd[2]));
    rgba[0] = c.x; rgba[1] = c.y; rgba[2] = c.z; rgba[3] = c.w;
}

}  // extern "C"
