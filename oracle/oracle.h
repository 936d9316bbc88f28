/* oracle.h — interface of the CPU parity oracle (TEST INFRASTRUCTURE ONLY; see oracle.c). */
#ifndef UVT_ORACLE_H
#define UVT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The world exactly as the reference shaders see it (assets/shaders/map.glsl:11-19):
 * SSBO 10 `chunks`, SSBO 9 `data`, image unit 6 `model` (256^3 RGBA8, x fastest). */
typedef struct orc_world {
    uint32_t dim;            /* MAP_DIMENSION */
    const uint32_t *chunks;  /* u32[(dim/8)^3] */
    const uint32_t *bricks;  /* u32[n][512]    */
    const uint32_t *atlas;   /* u32[256*256*256], texel = R | G<<8 | B<<16 | A<<24 */
} orc_world;

typedef struct orc_camera { /* camera.glsl:2-6 */
    float cam_pos[4];
    float cam_mat[16];
    float fov;
    float _pad[3];
} orc_camera;

/* traceEntities with the code behind its early return made live (map.glsl:203-248) and the primary-pass entity
 * composite that the reference keeps commented out (primary.comp.glsl:45-54): SURVEY section 8 row f3. */
#define ORC_MAX_ENTITIES 32
typedef struct orc_entities {
    uint32_t mode;        /* 0 = boxes (live behaviour, shadow pass only), 1 = models (dead code live + primary composite) */
    uint32_t n;           /* entities in pos[] */
    float pos[ORC_MAX_ENTITIES][3]; /* `positions[]` (map.glsl:173-179): low corner of each entity's box */
    const uint32_t *model;/* size^3 RGBA8 texels, x fastest, then y, then z; NULL = texels [0,8)^3 of the atlas image, which is
                           * what `imageLoad(model, ivec3(pos) & 7)` reads (map.glsl:218) */
    uint32_t size;        /* model edge in voxels: 8 (the literal `bounds`), 16 or 32 (chicken.vox, game.zig:114); box edge = size/8 blocks */
    uint32_t max_steps;   /* 64 (map.glsl:214) */
} orc_entities;

typedef struct orc_params {
    uint32_t map_dim;
    uint32_t primary_max_steps; /* 192 */
    uint32_t shadow_max_steps;  /* 48  */
    float epsilon;              /* 0.001 */
    uint32_t entities;          /* run traceEntities in the shadow pass */
    const struct orc_entities *ent; /* NULL: the reference as it runs (five literal boxes, early return at map.glsl:199) */
} orc_params;

/* HitInfo (map.glsl:62-70) plus everything the explicit hit buffer and the byte counters need. */
typedef struct orc_hit {
    uint32_t data;       /* HitInfo.data */
    float hit_pos[3];    /* HitInfo.hit_pos (sub-voxel units) */
    float normal[3];     /* HitInfo.normal */
    uint32_t p[3];       /* `pos` at the hit */
    uint32_t face;       /* 1..6, 0 = miss */
    uint32_t block;      /* block word at the hit */
    uint32_t trips;      /* loop trips executed */
    uint32_t exit_kind;  /* 0 hit, 1 cap, 2 left the map */
    uint32_t t_in, t_chunk, t_block;
    float distance;
} orc_hit;

/* same layout as uvt_hit (include/uvt.h) */
typedef struct orc_hit_rec {
    uint32_t px, py, pz;
    uint32_t block;
    uint32_t color;
    float distance;
    uint16_t trips;
    uint8_t face;
    uint8_t exit_kind;
} orc_hit_rec;

typedef struct orc_counters {
    uint64_t rays, t_in, t_chunk, t_block, hits, early_out;
} orc_counters;

void orc_sky_dome2(const float rd[3], float col[3]);
void orc_trace_map(const orc_world *w, const float origin[3], const float dir[3], int max_steps, orc_hit *out);
int  orc_trace_entities(const float o[3], const float d[3], float max_distance);
/* The general form: entity list from `e`; mode 1 also runs the sub-model DDA.  out->data != 0 on a hit; out->p = model voxel,
 * out->block = entity index, out->trips = model-loop trips, out->hit_pos in WORLD space (map.glsl:229). */
void orc_trace_entities_ex(const orc_world *w, const orc_entities *e, float epsilon, const float o[3], const float d[3],
                           float max_distance, orc_hit *out);
void orc_primary_ray(const orc_camera *cam, float tan_half_fov, uint32_t W, uint32_t H, uint32_t px, uint32_t py,
                     uint32_t map_dim, float epsilon, float origin[3], float dir[3], float start[3]);
void orc_primary(const orc_world *w, const orc_camera *cam, const orc_params *prm, uint32_t W, uint32_t H,
                 uint32_t *albedo, uint32_t *normal, float *position, orc_hit_rec *hits, orc_counters *counters);
void orc_primary_pixels(const orc_world *w, const orc_camera *cam, const orc_params *prm, uint32_t W, uint32_t H, size_t n,
                        const uint32_t *xs, const uint32_t *ys, uint32_t *albedo, uint32_t *normal, float *position, orc_hit_rec *hits);
void orc_secondary(const orc_world *w, const orc_params *prm, uint32_t W, uint32_t H,
                   const uint32_t *normal, const float *position, uint32_t *illum, orc_counters *counters);
void orc_blit(uint32_t W, uint32_t H, const uint32_t *albedo, const uint32_t *normal, const float *position,
              const uint32_t *illum, uint32_t *frame);
int  orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
