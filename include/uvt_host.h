/*
 * uvt_host.h — host-side data formats on either side of the traversal path, C ABI.
 *
 * These mirror the reference's Zig host types that own the memory the shaders read
 * (SURVEY.md §8 a13 / App. B).  They are CPU code with no CUDA dependency; a brickmap or
 * atlas can optionally be attached to a uvt_ctx, in which case its storage IS the pinned
 * staging of uvt_world_alloc / it forwards uploads to uvt_atlas_upload, exactly as the
 * reference types wrap GL buffers and textures.
 */
#ifndef UVT_HOST_H
#define UVT_HOST_H

#include "uvt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Voxel word: src/engine/voxel.zig:7-19 ---------------------------------------- */
#define UVT_VOXEL_TY_MASK 0x0FFFFFFFu
#define UVT_VOXEL_SOLID   0x10000000u
static inline uint32_t uvt_voxel(uint32_t ty, int is_solid) {
    return (ty & UVT_VOXEL_TY_MASK) | (is_solid ? UVT_VOXEL_SOLID : 0u);
}

/* ---- LCG: src/engine/util.zig:33-45 ------------------------------------------------ */
typedef struct uvt_lcg { uint32_t seed; } uvt_lcg;
static inline uint32_t uvt_lcg_rand(uvt_lcg *g) {
    g->seed = g->seed * 1103515245u + 12345u;  /* wrapping, returns the whole state */
    return g->seed;
}

/* ---- VoxelBrickmap(dim, 8): src/engine/voxel.zig:25-82 ----------------------------- */
typedef struct uvt_brickmap uvt_brickmap;
/* ctx may be NULL (plain host memory).  Initial pool capacity = dim bricks
 * (GpuBlockAllocator.init(dim), voxel.zig:36), doubling on exhaustion. */
int      uvt_brickmap_create(uvt_ctx *ctx, uint32_t dim, uvt_brickmap **out);
/* The same over a multi-GPU group (uvt.h, uvt_group): one staging, bind() publishes to every member. */
int      uvt_brickmap_create_group(uvt_group *group, uint32_t dim, uvt_brickmap **out);
void     uvt_brickmap_destroy(uvt_brickmap *bm);
void     uvt_brickmap_clear(uvt_brickmap *bm);                                /* voxel.zig:41-44 */
int      uvt_brickmap_set(uvt_brickmap *bm, uint32_t x, uint32_t y, uint32_t z, uint32_t voxel); /* :58-61 */
uint32_t uvt_brickmap_get(uvt_brickmap *bm, uint32_t x, uint32_t y, uint32_t z);                 /* :63-70 */
int      uvt_brickmap_is_walkable(uvt_brickmap *bm, uint32_t x, uint32_t y, uint32_t z);         /* :72-75 */
uint32_t uvt_brickmap_dim(const uvt_brickmap *bm);
size_t   uvt_brickmap_n_bricks(const uvt_brickmap *bm);      /* GpuBlockAllocator.block_index     */
size_t   uvt_brickmap_capacity(const uvt_brickmap *bm);      /* GpuBlockAllocator.max_block_index */
const uint32_t *uvt_brickmap_chunks(const uvt_brickmap *bm); /* u32[(dim/8)^3]      */
const uint32_t *uvt_brickmap_bricks(const uvt_brickmap *bm); /* u32[capacity][512]  */
/* VoxelBrickmap.bind (voxel.zig:77-80; called every frame, game.zig:236): publish to the attached ctx.
 * The first bind is a full uvt_world_commit; later binds publish only the box of blocks written through
 * uvt_brickmap_set since the previous bind (uvt_world_commit_region) and cost nothing when the map is clean.
 * Writes made through the raw pointers above are not tracked: call uvt_brickmap_mark_dirty after them. */
int      uvt_brickmap_bind(uvt_brickmap *bm);
void     uvt_brickmap_mark_dirty(uvt_brickmap *bm);
/* Reproducible world dump: "UVTW" u32 version, dim, n_bricks, chunks[], bricks[n][512]. */
int      uvt_brickmap_save(const uvt_brickmap *bm, const char *path);
int      uvt_brickmap_load(uvt_ctx *ctx, const char *path, uvt_brickmap **out);

/* ---- procgen: src/procgen.zig:6-70 -------------------------------------------------- */
int   uvt_procgen(uvt_brickmap *bm, uint32_t dim, float offset_x, float offset_y);
/* The same world generated on the device (uvt_world_procgen_plan / _fill, include/uvt.h): bm must be attached to a ctx and
 * empty.  UVT_ERR_INVALID when the device path does not apply (no ctx, a group, a non-empty map, an unsupported world):
 * call uvt_procgen then. */
int   uvt_procgen_device(uvt_brickmap *bm, uint32_t dim, float offset_x, float offset_y);
/* FastNoiseLite-style OpenSimplex2 FBm with znoise FnlGenerator defaults (procgen.zig:7);
 * restated from the published algorithm, NOT byte-verified against the library. */
float uvt_noise2_fbm(float x, float y);
/* terrain height of a column as procgen computes it (procgen.zig:23-24) */
uint32_t uvt_procgen_height(uint32_t dim, uint32_t x, uint32_t z, float offset_x, float offset_y);

/* ---- .vox reader (stands in for zvox.VoxFile.from_reader; voxel.zig:115-127) --------- */
typedef struct uvt_vox_voxel { uint8_t x, y, z, color; } uvt_vox_voxel;
typedef struct uvt_vox_file uvt_vox_file;
int      uvt_vox_parse(const void *bytes, size_t n, uvt_vox_file **out);
int      uvt_vox_open(const char *path, uvt_vox_file **out);
void     uvt_vox_free(uvt_vox_file *f);
uint32_t uvt_vox_n_models(const uvt_vox_file *f);
int      uvt_vox_model_size(const uvt_vox_file *f, uint32_t model, uint32_t size_xyz[3]);
uint32_t uvt_vox_model_n_voxels(const uvt_vox_file *f, uint32_t model);
const uvt_vox_voxel *uvt_vox_model_voxels(const uvt_vox_file *f, uint32_t model);
const uint32_t *uvt_vox_palette(const uvt_vox_file *f); /* 256 raw RGBA entries, entry k <-> colour index k+1 */
const char *uvt_vox_error(void);

/* ---- VoxelModelAtlas: src/engine/voxel.zig:84-132 ------------------------------------ */
typedef struct uvt_atlas uvt_atlas;
int      uvt_atlas_create(uvt_ctx *ctx, uvt_atlas **out);       /* 256^3 RGBA8 (voxel.zig:88-92); ctx may be NULL */
int      uvt_atlas_create_group(uvt_group *group, uvt_atlas **out);   /* models are uploaded to every member */
void     uvt_atlas_destroy(uvt_atlas *a);
/* load_block_model (voxel.zig:115-127): every model of the file → next slots, y/z swapped. */
int      uvt_atlas_load_block_model(uvt_atlas *a, const char *path);
int      uvt_atlas_load_block_model_mem(uvt_atlas *a, const void *bytes, size_t n);
/* append one prebuilt 8^3 model (512 texels, x + 8*y + 64*z) — used with fixture tables */
int      uvt_atlas_append_model(uvt_atlas *a, const uint32_t texels[512]);
uint32_t uvt_atlas_current_index(const uvt_atlas *a);
/* the 512 texels of slot `idx` */
int      uvt_atlas_get_model(const uvt_atlas *a, uint32_t idx, uint32_t texels[512]);

/* ---- Camera: src/engine/graphics/camera.zig:5-43 -------------------------------------- */
typedef struct uvt_camera_state {
    float fov, pitch, yaw;
    float cam_mat[16];
    float cam_pos[4];
} uvt_camera_state;
void uvt_camera_init(uvt_camera_state *c);                            /* fov=pi/2, identity */
void uvt_camera_rotate(uvt_camera_state *c, float pitch, float yaw);  /* camera.zig:18-23  */
void uvt_camera_set_pos(uvt_camera_state *c, const float pos[4]);
void uvt_camera_increment_fov(uvt_camera_state *c, float increment);  /* camera.zig:29-31  */
void uvt_camera_as_uniform_data(const uvt_camera_state *c, uvt_camera *out); /* :37-43 */
/* zmath.matFromRollPitchYaw(pitch, yaw, 0) restated (row vectors, v' = v*M). */
void uvt_mat_from_pitch_yaw(float pitch, float yaw, float out16[16]);

#ifdef __cplusplus
}
#endif
#endif /* UVT_HOST_H */
