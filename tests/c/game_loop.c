/*
 * game_loop.c — the reference's renderer call order (src/game.zig:54-129 game_init, :224-229 pre_render,
 * :232-256 render) driven through the C ABI from plain C, the way the Zig glue of INTEGRATION.md would.
 *
 *   game_loop <atlas_models.bin> <dim> <width> <height> <out_frame.bin>
 *
 * atlas_models.bin: n * 512 little-endian u32 texels (x + 8y + 64z per model), i.e. tests/golden/atlas_models.npy
 * without its header.  Exit codes: 0 ok, 3 no CUDA device (UVT_ERR_NO_DEVICE: there is no CPU path), 1 any other error.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "uvt.h"
#include "uvt_host.h"

#define CHECK(ctx, call)                                                                  \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != UVT_OK) {                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, uvt_last_error(ctx));           \
            return rc_ == UVT_ERR_NO_DEVICE ? 3 : 1;                                      \
        }                                                                                 \
    } while (0)

int main(int argc, char **argv) {
    if (argc != 6) {
        fprintf(stderr, "usage: %s atlas_models.bin dim width height out_frame.bin\n", argv[0]);
        return 1;
    }
    const uint32_t dim = (uint32_t)atoi(argv[2]), width = (uint32_t)atoi(argv[3]), height = (uint32_t)atoi(argv[4]);

    /* gfx.init (opengl.zig:25-34 -> graphics.zig:60) */
    uvt_params params;
    uvt_default_params(&params);
    params.map_dim = dim;
    uvt_ctx *ctx = NULL;
    CHECK(NULL, uvt_create(&params, 0, &ctx));

    /* pipelines (game.zig:79-89) */
    uvt_pipeline *primary = NULL, *secondary = NULL, *edit = NULL, *raster = NULL;
    CHECK(ctx, uvt_pipeline_create(ctx, UVT_PIPELINE_PRIMARY, &primary));
    CHECK(ctx, uvt_pipeline_create(ctx, UVT_PIPELINE_SECONDARY, &secondary));
    CHECK(ctx, uvt_pipeline_create(ctx, UVT_PIPELINE_EDIT, &edit));
    CHECK(ctx, uvt_pipeline_create(ctx, UVT_PIPELINE_BLIT, &raster));

    /* GBuffer.init (game.zig:91) */
    CHECK(ctx, uvt_resize(ctx, width, height));

    /* VoxelBrickmap.init + procgen (game.zig:96-97): the brickmap writes straight into the ctx's pinned staging */
    uvt_brickmap *voxels = NULL;
    CHECK(ctx, uvt_brickmap_create(ctx, dim, &voxels));
    CHECK(ctx, uvt_procgen(voxels, dim, 0.0f, 0.0f));

    /* VoxelModelAtlas.init + load_block_model x13 (game.zig:99-113): here from the prebuilt model table */
    uvt_atlas *models = NULL;
    CHECK(ctx, uvt_atlas_create(ctx, &models));
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    uint32_t texels[512];
    while (fread(texels, sizeof texels, 1, f) == 1) CHECK(ctx, uvt_atlas_append_model(models, texels));
    fclose(f);

    /* update (game.zig:207-212): camera = player position + (0, 3, 0) */
    uvt_camera_state cam;
    uvt_camera_init(&cam);
    const float pos[4] = {dim / 2.0f, 22.0f + 3.0f, dim / 2.0f, 0.0f};
    uvt_camera_set_pos(&cam, pos);

    for (int frame = 0; frame < 3; ++frame) {
        /* pre_render (game.zig:224-229) */
        uvt_camera ubo;
        uvt_camera_as_uniform_data(&cam, &ubo);
        /* render (game.zig:232-256) */
        CHECK(ctx, uvt_set_camera(ctx, &ubo));                                  /* cam_uniforms.bind(8)  */
        CHECK(ctx, uvt_brickmap_bind(voxels));                                  /* voxels.bind(9)        */
        const uint32_t gx = width / 32 + 1, gy = height / 32 + 1;               /* game.zig:241-242      */
        CHECK(ctx, uvt_pipeline_dispatch(primary, gx, gy, 1));
        CHECK(ctx, uvt_pipeline_dispatch(secondary, gx, gy, 1));
        CHECK(ctx, uvt_pipeline_dispatch(raster, 1, 1, 1));                     /* raster.draw(4)        */
        uvt_camera_rotate(&cam, 40.0f, 250.0f);                                 /* mouse_moved           */
    }
    /* the frame rendered BEFORE the last rotate: re-render frame 0's camera for a deterministic output */
    uvt_camera_init(&cam);
    uvt_camera_set_pos(&cam, pos);
    uvt_camera ubo;
    uvt_camera_as_uniform_data(&cam, &ubo);
    CHECK(ctx, uvt_set_camera(ctx, &ubo));
    CHECK(ctx, uvt_dispatch_frame(ctx));

    const size_t bytes = uvt_buffer_bytes(ctx, UVT_BUF_FRAME);
    void *frame_host = malloc(bytes);
    CHECK(ctx, uvt_readback(ctx, UVT_BUF_FRAME, frame_host, bytes));
    uvt_hit pick;
    CHECK(ctx, uvt_pick(ctx, &pick));                                           /* terrain_edit pick ray */
    f = fopen(argv[5], "wb");
    if (!f || fwrite(frame_host, 1, bytes, f) != bytes) { perror(argv[5]); return 1; }
    fclose(f);
    printf("frame %ux%u, %zu bytes, pick face %u trips %u, %llu kernel launches\n", width, height, bytes, (unsigned)pick.face,
           (unsigned)pick.trips, (unsigned long long)uvt_launch_count(ctx));

    free(frame_host);
    uvt_pipeline_destroy(primary); uvt_pipeline_destroy(secondary); uvt_pipeline_destroy(edit); uvt_pipeline_destroy(raster);
    uvt_atlas_destroy(models);
    uvt_brickmap_destroy(voxels);
    uvt_destroy(ctx);
    return 0;
}
