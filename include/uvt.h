/*
 * uvt.h — C ABI of the B200-native voxel ray-traversal pass.
 *
 * This is the drop-in boundary for the reference engine's renderer entry points
 * (SURVEY.md §8b).  Every export below names the reference interface it replaces
 * (paths relative to the reference checkout).  Plain pointers and sizes only; no
 * torch / C++ types cross this boundary.
 *
 * Conventions
 *   - every call returns 0 on success or a negative uvt_status; the message is
 *     available through uvt_last_error().  Nothing aborts (the reference @panic's).
 *   - one ctx per host thread; all device work of a ctx is issued on ONE CUDA stream
 *     in call order (the reference uses one in-order GL queue + glMemoryBarrier,
 *     src/engine/graphics/shader.zig:113-117).  Dispatches are asynchronous like
 *     glDispatchCompute; uvt_sync()/uvt_readback() wait.
 *   - there is no CPU fallback: without a CUDA device uvt_create() fails.
 */
#ifndef UVT_H
#define UVT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVT_ABI_VERSION 2

typedef enum uvt_status {
    UVT_OK = 0,
    UVT_ERR_INVALID = -1,  /* bad argument / call order                      */
    UVT_ERR_CUDA = -2,     /* CUDA runtime error (message has the detail)    */
    UVT_ERR_NO_DEVICE = -3,/* no usable CUDA device: there is no CPU path    */
    UVT_ERR_OOM = -4,
    UVT_ERR_FORMAT = -5,   /* malformed .vox / world dump                    */
    UVT_ERR_IO = -6
} uvt_status;

/* ---- parameters: every constant the reference hard-codes in source ------------- */
typedef struct uvt_params {
    uint32_t map_dim;            /* MAP_DIMENSION, blocks per axis  (assets/shaders/map.glsl:2; src/game.zig:33) */
    uint32_t primary_max_steps;  /* 192  (assets/shaders/primary.comp.glsl:43)   */
    uint32_t shadow_max_steps;   /* 48   (assets/shaders/secondary.comp.glsl:41) */
    uint32_t edit_max_steps;     /* 64   (assets/shaders/terrain_edit.comp.glsl:16) */
    float    epsilon;            /* EPSILON 0.001 (assets/shaders/map.glsl:7)    */
    uint32_t flags;              /* UVT_FLAG_*                                    */
    uint32_t layout;             /* uvt_layout: which device world layout the traversal kernels read */
    uint32_t scheduler;          /* uvt_scheduler: how rays are assigned to lanes (B200 layout only) */
    uint32_t reserved[8];
} uvt_params;

enum {
    UVT_FLAG_HIT_BUFFER = 1u << 0, /* also write the explicit hit buffer (uvt_hit) in the primary pass */
    UVT_FLAG_ENTITIES   = 1u << 1, /* shadow pass runs traceEntities (map.glsl:172-201); on by default */
    UVT_FLAG_NO_DENSE   = 1u << 2, /* do not use the dense block grid (dim^3 bytes); traverse through chunk table + bricks */
    UVT_FLAG_FUSED_FRAME = 1u << 3 /* uvt_dispatch_frame runs ONE fused kernel instead of the (faster, default) three launches */
};

typedef enum uvt_layout {
    UVT_LAYOUT_COMPACT   = 0, /* B200 layout: chunk-occupancy window + 8-bit material bricks + model bitmasks */
    UVT_LAYOUT_REFERENCE = 1  /* the reference SSBO layout read verbatim (u32 chunk table + u32[512] bricks + RGBA8 atlas) */
} uvt_layout;

typedef enum uvt_scheduler {
    UVT_SCHED_TILE = 0, /* one pixel per thread for the whole traversal (default — also what a zero-initialised uvt_params selects) */
    UVT_SCHED_POOL = 1  /* per-CTA ray pool, compacted between trip phases (opt-in; measured slower) */
} uvt_scheduler;

void uvt_default_params(uvt_params *p);

/* ---- camera: src/engine/graphics/camera.zig:12-16 == camera.glsl:2-6 (std140) --- */
typedef struct uvt_camera {
    float cam_pos[4];   /* C_position (w unused)                        @0  */
    float cam_mat[16];  /* C_view: 4 rows of F32x4 = 4 GLSL columns     @16 */
    float fov;          /* radians                                      @80 */
    float _pad[3];      /* std140 tail                                  @84 */
} uvt_camera;           /* 96 bytes */

/* ---- explicit hit buffer (new; the reference never materialises it) */
typedef struct uvt_hit {
    uint32_t px, py, pz; /* hit sub-voxel coordinate `pos` (map.glsl:108); 0xFFFFFFFF on miss */
    uint32_t block;      /* block word: ty | is_solid<<28 (src/engine/voxel.zig:7-19); 0 on miss */
    uint32_t color;      /* atlas texel R|G<<8|B<<16|A<<24 (= HitInfo.data); 0 on miss          */
    float    distance;   /* length(hit_pos/8 - rayOrigin) in blocks, fp32; -1 on miss            */
    uint16_t trips;      /* DDA loop trips executed (map.glsl:106)                               */
    uint8_t  face;       /* faceId 1..6 (map.glsl:119-125), 0 = miss                             */
    uint8_t  exit_kind;  /* 0 hit, 1 step cap exhausted, 2 left the map, 3 entity (UVT_ENTITY_MODELS composite:
                          * p = model voxel, block = 0x80000000 | entity index, trips = model-loop trips) */
} uvt_hit;               /* 28 bytes */

/* Exact per-pass traversal counters; they define the ALGORITHMIC bytes of SURVEY §8d:
 * bytes = 4*t_in + 4*t_chunk + 4*t_block + per-pixel G-buffer traffic. */
typedef struct uvt_counters {
    uint64_t rays;       /* rays traced (pixels that ran traceMap)                         */
    uint64_t t_in;       /* loop trips that passed the bounds test (one chunks[] read)     */
    uint64_t t_chunk;    /* trips whose chunk entry != 0 (one data[] read)                 */
    uint64_t t_block;    /* trips whose block != 0 (one atlas read)                        */
    uint64_t hits;       /* rays that returned a hit                                       */
    uint64_t early_out;  /* secondary only: pixels that left at secondary.comp.glsl:26-29  */
} uvt_counters;

typedef struct uvt_ctx uvt_ctx;
typedef struct uvt_pipeline uvt_pipeline;

/* ---- context: gfx.init / enableDebug (src/engine/graphics/graphics.zig:43-75) ---- */
int  uvt_create(const uvt_params *params, int device, uvt_ctx **out);
void uvt_destroy(uvt_ctx *ctx);
/* ctx may be NULL: returns the calling thread's last creation-time error. */
const char *uvt_last_error(uvt_ctx *ctx);
int  uvt_abi_version(void);
/* Issue all further work on an externally owned cudaStream_t (NULL restores the ctx's own stream). */
int  uvt_set_stream(uvt_ctx *ctx, void *cuda_stream);
int  uvt_get_params(uvt_ctx *ctx, uvt_params *out);
/* Switch the device world layout the traversal kernels read (both stay resident after a commit). */
int  uvt_set_layout(uvt_ctx *ctx, uint32_t layout);
int  uvt_set_scheduler(uvt_ctx *ctx, uint32_t scheduler);
/* The layout actually in use: COMPACT needs <= 255 distinct block words, else REFERENCE is used. */
int  uvt_effective_layout(uvt_ctx *ctx);
/* Step caps (reference constants 192 / 48: primary.comp.glsl:43, secondary.comp.glsl:41). */
int  uvt_set_max_steps(uvt_ctx *ctx, uint32_t primary, uint32_t shadow);

/* ---- pipelines: ComputePipeline / RasterPipeline init+deinit (shader.zig:97-153) --
 * Kernels are precompiled, so a pipeline is an opaque token naming a pass; creating one
 * checks that the kernel image for this device is loadable (the analogue of compile+link),
 * and hot-reload (src/game.zig:258-288) is create-new + destroy-old. */
typedef enum uvt_pipeline_kind {
    UVT_PIPELINE_PRIMARY = 0,   /* assets/shaders/primary.comp.glsl   */
    UVT_PIPELINE_SECONDARY = 1, /* assets/shaders/secondary.comp.glsl */
    UVT_PIPELINE_EDIT = 2,      /* assets/shaders/terrain_edit.comp.glsl */
    UVT_PIPELINE_BLIT = 3       /* assets/shaders/blit.{vertex,fragment}.glsl */
} uvt_pipeline_kind;
int  uvt_pipeline_create(uvt_ctx *ctx, uvt_pipeline_kind kind, uvt_pipeline **out);
void uvt_pipeline_destroy(uvt_pipeline *p);
/* ComputePipeline.dispatch(x,y,z) (shader.zig:113-117): group counts are accepted for
 * signature parity and validated against the G-buffer ((W/32+1)x(H/32+1), game.zig:241-242),
 * the pass always covers the whole G-buffer.  RasterPipeline.draw(4) maps to kind BLIT. */
int  uvt_pipeline_dispatch(uvt_pipeline *p, uint32_t gx, uint32_t gy, uint32_t gz);

/* ---- world: VoxelBrickmap + GpuBlockAllocator (src/engine/voxel.zig:25-82,
 *      src/engine/graphics/gpu_block_allocator.zig:4-41, buffer.zig:80-134) ---------
 * The reference host writes chunk table and brick pool in place through persistent
 * mappings.  Here the ctx hands out PINNED host staging with the same layout
 * (chunks: u32[(dim/8)^3], 0 = empty else brick+1; bricks: u32[capacity][512],
 * index x%8 + 8*(y%8) + 64*(z%8)); uvt_world_commit() uploads and repacks it. */
int  uvt_world_alloc(uvt_ctx *ctx, uint32_t dim, uint32_t **chunks_host, uint32_t **bricks_host, size_t brick_capacity);
/* The same with CALLER-OWNED staging (the host keeps its own chunk table and brick pool, same layout; pinned memory
 * uploads asynchronously, pageable memory works too).  dim = 0 afterwards re-points the brick pool after the caller
 * grew or moved it (chunks_host is ignored).  The ctx never frees these; uvt_world_grow is refused. */
int  uvt_world_use_staging(uvt_ctx *ctx, uint32_t dim, uint32_t *chunks_host, uint32_t *bricks_host, size_t brick_capacity);
/* GpuBlockAllocator.alloc growth (gpu_block_allocator.zig:20-24 → buffer.zig:48-62): contents are preserved. */
int  uvt_world_grow(uvt_ctx *ctx, size_t new_capacity, uint32_t **bricks_host);
/* Make host edits visible to the GPU (H2D + repack kernel). n_bricks = GpuBlockAllocator.block_index. */
int  uvt_world_commit(uvt_ctx *ctx, size_t n_bricks);
/* Incremental publish: same result as uvt_world_commit(n_bricks) when the staging differs from
 * the last committed state only inside the block box [lo, hi] (inclusive block coordinates, hi < dim) — the
 * live-mapping semantics of VoxelBrickmap.set (voxel.zig:58-64) at interactive cost.  Existing chunk entries
 * must be unchanged; new bricks (indices >= the previous n_bricks) may be attached to empty chunks of the box.
 * Falls back to a full commit by itself whenever the in-place path does not apply (first commit, box of more
 * than 4096 chunks, no spare brick slots, more than 223 materials). */
int  uvt_world_commit_region(uvt_ctx *ctx, size_t n_bricks, const uint32_t lo[3], const uint32_t hi[3]);
/* map_setVoxel (map.glsl:49-55): writes the block word into the staging AND publishes it, but only where the
 * chunk already holds a brick (the shader never allocates); *written (may be NULL) tells which. */
int  uvt_world_set_voxel(uvt_ctx *ctx, uint32_t x, uint32_t y, uint32_t z, uint32_t voxel, int *written);
/* Checksums of the layout the commit derived (test hook: an incremental commit must leave exactly what a full
 * commit builds).  Keyed by block position, independent of brick numbering: out[0] dense grid, out[1] brick
 * view (brick bytes + far-empty chunk entries), out[2] column-group tops + per-column tops and sun clearance, out[3] = y_clear |
 * n_materials << 32. */
int  uvt_world_layout_checksum(uvt_ctx *ctx, uint64_t out[4]);

/* procgen on the device (src/procgen.zig:6-70): the world the serial host procgen (uvt_procgen, uvt_host.h)
 * writes — the same chunk table, brick numbering and brick words, byte for byte — generated by kernels.  Two steps because
 * the brick count is only known after the plan: _plan returns it, the caller grows the staging the way GpuBlockAllocator
 * would (uvt_world_grow), _fill writes device AND host staging.  uvt_world_commit(n_bricks) publishes it as usual.
 * UVT_ERR_INVALID with "unsupported world" = more overlapping trees than the device path tracks: use the host procgen. */
int  uvt_world_procgen_plan(uvt_ctx *ctx, float offset_x, float offset_y, size_t *n_bricks);
int  uvt_world_procgen_fill(uvt_ctx *ctx);

/* ---- atlas: VoxelModelAtlas → Texture.set_data_offset → glTextureSubImage3D
 *      (voxel.zig:88-131, texture.zig:70-72): RGBA8 sub-box, x fastest then y then z. */
int  uvt_atlas_upload(uvt_ctx *ctx, uint32_t ox, uint32_t oy, uint32_t oz,
                      uint32_t w, uint32_t h, uint32_t d, const uint32_t *rgba);

/* ---- camera UBO: PersistentMappedBuffer(UniformData) (game.zig:94,224-229,235) ---- */
int  uvt_set_camera(uvt_ctx *ctx, const uvt_camera *cam);
/* Batched poses (BASELINE config 5): pose i renders into G-buffer layer i. */
int  uvt_set_cameras(uvt_ctx *ctx, const uvt_camera *cams, int n);

/* ---- G-buffer: GBuffer.init/resize (gbuffer.zig:9-30; game.zig:91,197-205) -------- */
int  uvt_resize(uvt_ctx *ctx, uint32_t width, uint32_t height);
/* Multi-GPU image partition: this ctx renders only the row bands b with
 * b % n_parts == part (band = band_rows image rows) of the full WxH frame, stored
 * compactly band after band.  n_parts = 1 restores whole-frame rendering. */
int  uvt_set_partition(uvt_ctx *ctx, uint32_t band_rows, uint32_t n_parts, uint32_t part);
/* rows this ctx renders under the current partition */
int  uvt_local_rows(uvt_ctx *ctx, uint32_t *rows);

/* ---- passes: game.zig:244-255 ------------------------------------------------------ */
int  uvt_dispatch_primary(uvt_ctx *ctx);    /* primary.comp.glsl main   */
int  uvt_dispatch_secondary(uvt_ctx *ctx);  /* secondary.comp.glsl main */
int  uvt_shade(uvt_ctx *ctx);               /* blit.fragment.glsl main → RGBA8 frame */
/* the two calls above in ONE launch (what uvt_dispatch_frame runs after the primary pass): same illumination image, same frame */
int  uvt_dispatch_secondary_shade(uvt_ctx *ctx);
/* primary+secondary+shade in one launch; results identical to the three calls above */
int  uvt_dispatch_frame(uvt_ctx *ctx);
/* ---- entities: traceEntities (assets/shaders/map.glsl:172-248), SURVEY 8 row f3 ----------------------
 * The reference as it runs leaves traceEntities at map.glsl:199 (five literal unit boxes intersected as lines, shadow
 * pass only): UVT_ENTITY_BOXES, the default, compiled into the shadow kernel.  UVT_ENTITY_MODELS makes the code
 * behind that return live — a DDA over the entity's voxel model (map.glsl:203-248) — together with the entity
 * composite the reference keeps commented out in the primary pass (primary.comp.glsl:45-54); both then run as
 * extra launches after the primary / secondary kernels.  MODELS turns the hit buffer on (the composite reads the
 * terrain distance of primary.comp.glsl:47 from it): call it before uvt_resize, or expect the G-buffer to be
 * re-created.  UVT_FLAG_ENTITIES off disables every entity test. */
typedef enum uvt_entity_mode {
    UVT_ENTITY_BOXES = 0,
    UVT_ENTITY_MODELS = 1
} uvt_entity_mode;
int  uvt_set_entity_mode(uvt_ctx *ctx, uint32_t mode);
/* `positions[]` of map.glsl:173-179: n <= 32 low box corners (xyz, blocks); n = 0 restores the five literals. */
int  uvt_set_entities(uvt_ctx *ctx, const float *positions_xyz, uint32_t n);
/* The entity model: size^3 RGBA8 texels (size 8, 16 or 32: chicken.vox is 32^3, src/game.zig:114), x fastest, then
 * y, then z; the entity's box edge is size/8 blocks, so a model voxel is as large as a world sub-voxel.  rgba = NULL
 * restores the text as written: `imageLoad(model, ivec3(pos) & 7)` = texels [0,8)^3 of the atlas (map.glsl:218).
 * max_steps = the cap of the model loop (map.glsl:214), 0 = 64. */
int  uvt_entity_model_upload(uvt_ctx *ctx, uint32_t size, const uint32_t *rgba, uint32_t max_steps);

/* uvt_dispatch_frame in n row chunks (1 = whole-frame launches, the default): chunk k runs its three passes on the ctx
 * stream, chunk k+1 on a side stream, so the tail of one pass (a few long-running warps) overlaps the next chunk's work.
 * Worth it when a GPU's share of the frame is small (tiled frames at N = 4, 8); identical pixels.  With n > 1 only
 * the frame has a time (uvt_last_pass_ms 3). */
int  uvt_set_frame_chunks(uvt_ctx *ctx, uint32_t n);
/* terrain_edit.comp.glsl: the centre pick ray, traceMap(...,64); returns the hit. */
int  uvt_pick(uvt_ctx *ctx, uvt_hit *out);
int  uvt_sync(uvt_ctx *ctx);

typedef enum uvt_buffer_kind {
    UVT_BUF_ALBEDO = 0,       /* RGBA8   4 B/px  (image unit 0) */
    UVT_BUF_NORMAL = 1,       /* RGBA8   4 B/px  (image unit 1) */
    UVT_BUF_POSITION = 2,     /* RGBA32F 16 B/px (image unit 2) */
    UVT_BUF_ILLUMINATION = 3, /* RGBA8   4 B/px  (image unit 3) */
    UVT_BUF_FRAME = 4,        /* RGBA8   4 B/px  default framebuffer after the blit */
    UVT_BUF_HIT = 5           /* uvt_hit 28 B/px (needs UVT_FLAG_HIT_BUFFER) */
} uvt_buffer_kind;
/* Row 0 is the BOTTOM image row (GL image origin). Copies layer 0..n_layers-1 back to back. */
int  uvt_readback(uvt_ctx *ctx, uvt_buffer_kind kind, void *dst, size_t bytes);
size_t uvt_buffer_bytes(uvt_ctx *ctx, uvt_buffer_kind kind);
/* Pipelined readback (what a presenting loop does instead of a blocking glReadPixels): the buffer is snapshotted on the
 * ctx stream (device-to-device, so the next dispatch may overwrite it at once) and copied to `dst` (pinned host memory
 * from uvt_alloc_pinned) on a separate copy stream while the next frame renders.  At most two readbacks are in flight;
 * a third call first waits for the oldest.  uvt_readback_wait() returns when every outstanding copy has landed. */
int  uvt_readback_async(uvt_ctx *ctx, uvt_buffer_kind kind, void *dst, size_t bytes);
int  uvt_readback_wait(uvt_ctx *ctx);
/* Tiled frames end to end: copy this ctx's FRAME bands to where they belong in a full W x H RGBA8 host
 * frame.  Every rank calls it on the same host frame (shared memory, page-locked with uvt_host_register in each
 * process): the frame is assembled in host memory by N parallel device-to-host copies, one per PCIe link, instead of
 * being gathered onto one GPU first.  Pipelined like uvt_readback_async; uvt_readback_wait() waits for it. */
int  uvt_readback_bands_async(uvt_ctx *ctx, void *host_frame, size_t frame_bytes);
int  uvt_host_register(uvt_ctx *ctx, void *p, size_t bytes);
int  uvt_host_unregister(uvt_ctx *ctx, void *p);
/* Device address of a buffer (for NCCL / peer access by the caller's communication layer). */
int  uvt_device_ptr(uvt_ctx *ctx, uvt_buffer_kind kind, void **dptr);
/* Redirect the FRAME output to caller-owned device memory (may be a peer-mapped
 * pointer of another GPU: finished bands are then stored straight over NVLink). */
int  uvt_bind_frame_target(uvt_ctx *ctx, void *dptr, uint32_t reserved, uint32_t global_rows);
/* Peer-to-peer frame (the fused alternative to the gather): the presenting rank creates a full WxH RGBA8 frame
 * and exports a CUDA IPC handle (64 bytes); every other rank opens it and binds the mapped pointer with
 * uvt_bind_frame_target(ptr, 0, 1): its kernels then store finished pixels straight into the presenting
 * GPU's memory over NVLink, so no gather and no reassembly pass remain. */
int  uvt_shared_frame_create(uvt_ctx *ctx, void **dptr, unsigned char handle_out[64]);
int  uvt_shared_frame_open(uvt_ctx *ctx, const unsigned char handle[64], void **dptr);
int  uvt_shared_frame_close(uvt_ctx *ctx, void *dptr);
/* Copy `bytes` from any device pointer visible to this ctx (e.g. the shared frame) to host memory. */
int  uvt_read_device(uvt_ctx *ctx, const void *dptr, void *dst, size_t bytes);
/* Rank 0 after the NCCL gather: `gathered` holds n_parts compact band buffers of rows_per_part rows
 * back to back; writes the assembled WxH frame (one kernel on the ctx stream). */
int  uvt_deinterleave(uvt_ctx *ctx, const void *gathered, void *frame, uint32_t rows_per_part);
/* Pinned host memory for readback targets / upload sources (the e2e path of bench.py). */
int  uvt_alloc_pinned(uvt_ctx *ctx, size_t bytes, void **out);
int  uvt_free_pinned(uvt_ctx *ctx, void *p);

/* Exact traversal counters of the last primary (which=0) / secondary (which=1) pass run
 * with counting on; uvt_count_pass re-runs that pass with the counting kernel variant. */
int  uvt_count_pass(uvt_ctx *ctx, int which, uvt_counters *out);
/* Device time of the last dispatch of each pass, measured with CUDA events on the ctx stream. */
int  uvt_last_pass_ms(uvt_ctx *ctx, int which /*0 primary,1 secondary,2 shade,3 frame*/, float *ms);
int  uvt_enable_timing(uvt_ctx *ctx, int on);
/* Number of kernels this ctx has launched since creation. */
uint64_t uvt_launch_count(uvt_ctx *ctx);

/* Utility kernel for measuring the L2 read bandwidth roofline denominator:
 * repeatedly reads an L2-resident buffer of `bytes` with 16-B loads; returns GB/s. */
int  uvt_measure_l2_read_gbps(uvt_ctx *ctx, size_t bytes, int repeats, float *gbps);
int  uvt_measure_hbm_copy_gbps(uvt_ctx *ctx, size_t bytes, int repeats, float *gbps);

/* ---- NCCL band exchange ("finished tiles are gathered to the presenting rank with NCCL over NVLink") ----
 * One ctx per rank (one process per GPU, or several ctxs in one process), partitioned with uvt_set_partition(band,
 * n_ranks, rank).  uvt_nccl_unique_id() is called once (rank 0) and its 128 bytes handed to every rank by the host's own
 * means; uvt_nccl_init() is collective.  uvt_dispatch_frame_nccl() renders the rank's bands in `n_groups` band groups and
 * exchanges each finished group with grouped ncclSend / ncclRecv on a second stream while the next group is traversed:
 * rank 0 receives every band straight at its rows of `full_frame` (device memory, W*H*4 bytes; NULL on the other ranks)
 * and shades its own bands into it, so the assembled frame needs no reassembly pass.  Stream-ordered like every dispatch:
 * the frame is complete on the ctx stream when the call's work is.  The library binds libnccl.so.2 at run time (in a
 * torchrun process: the copy torch loaded); without it these calls fail with a message, nothing else is affected.
 * The peer-to-peer band stores (uvt_shared_frame_*, uvt_group) remain the faster default; this is the NCCL variant. */
int  uvt_nccl_unique_id(unsigned char id_out[128]);
int  uvt_nccl_init(uvt_ctx *ctx, const unsigned char id[128], int n_ranks, int rank);
int  uvt_nccl_shutdown(uvt_ctx *ctx);
int  uvt_dispatch_frame_nccl(uvt_ctx *ctx, void *full_frame, uint32_t n_groups);

/* ---- several GPUs in ONE process (the reference is a single-process game, src/game.zig) ----------
 * A group owns one ctx per device.  World and atlas are replicated from one pinned staging; the frame is cut into
 * interleaved 16-row bands (member i renders bands i, i+n, ...) and every member's kernels store their finished
 * bands straight into the frame of member 0 over NVLink (peer access) — no gather pass.  Calls mirror the ctx calls
 * and fan out to every member; uvt_group_member(g, 0) presents (uvt_pick, timing, ...).  With n = 1 the group is a
 * plain ctx.  Only the FRAME is assembled; the G-buffers stay partitioned on their devices. */
typedef struct uvt_group uvt_group;
int  uvt_group_create(const uvt_params *params, const int *devices, int n, uvt_group **out);
void uvt_group_destroy(uvt_group *g);
int  uvt_group_size(const uvt_group *g);
uvt_ctx *uvt_group_member(uvt_group *g, int i);
const char *uvt_group_last_error(uvt_group *g);        /* g may be NULL after a failed create */
int  uvt_group_world_alloc(uvt_group *g, uint32_t dim, uint32_t **chunks_host, uint32_t **bricks_host, size_t brick_capacity);
int  uvt_group_world_grow(uvt_group *g, size_t new_capacity, uint32_t **bricks_host);
int  uvt_group_world_commit(uvt_group *g, size_t n_bricks);
int  uvt_group_world_commit_region(uvt_group *g, size_t n_bricks, const uint32_t lo[3], const uint32_t hi[3]);
int  uvt_group_atlas_upload(uvt_group *g, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t w, uint32_t h, uint32_t d, const uint32_t *rgba);
int  uvt_group_set_entity_mode(uvt_group *g, uint32_t mode);
int  uvt_group_set_entities(uvt_group *g, const float *positions_xyz, uint32_t n);
int  uvt_group_entity_model_upload(uvt_group *g, uint32_t size, const uint32_t *rgba, uint32_t max_steps);
int  uvt_group_set_camera(uvt_group *g, const uvt_camera *cam);
int  uvt_group_resize(uvt_group *g, uint32_t width, uint32_t height);
int  uvt_group_dispatch_frame(uvt_group *g);           /* primary + secondary + shade on every member, asynchronous */
int  uvt_group_sync(uvt_group *g);
int  uvt_group_readback_frame(uvt_group *g, void *dst, size_t bytes);   /* waits for every member, W*H*4 bytes */
int  uvt_group_frame_ptr(uvt_group *g, void **dptr);   /* the assembled frame on member 0's device (interop) */
int  uvt_group_count_pass(uvt_group *g, int which, uvt_counters *sum);  /* uvt_count_pass summed over the members */

#ifdef __cplusplus
}
#endif
#endif /* UVT_H */
