// procgen.cuh — src/procgen.zig:6-70 on the device, bit-identical to the serial host version (csrc/host/world.cpp:
// uvt_procgen): same chunk table, same brick numbering (first-touch order of VoxelBrickmap.set), same brick words.
//
// What is serial in the reference is ONE thing: the single LCG stream (util.zig:33-45) whose draw count per column depends
// on the values drawn (grass? flower? tree?) and, through the `continue` at procgen.zig:37-38, on trees planted by earlier
// columns.  Everything else is a function of (column, LCG state at the start of the column):
//   1. heights_kernel        vh[x][z] = u32(max(noise2 * dim * 0.1, 0))                         (parallel)
//   2. scan_kernel           ONE thread walks the columns in the reference order carrying the LCG state; the vh (+ up to
//                            16 sand, + 1 top layer) draws of a column body are skipped with the closed-form jump
//                            L^n(s) = A_n s + C_n (table), the 3-4 decoration draws are evaluated, trees are recorded and
//                            the columns they occupy remembered for the `continue` test.  ~26 dependent cycles per column.
//   3. touch_kernel/tree_touch_kernel   first set() that touches every chunk (64-bit key: x, z, event) -> radix sort ->
//                            brick numbers in the allocator's first-touch order (the y < 16 slab comes first, x, z, y order)
//   4. fill_kernel           every column replays its own draws from its start state and writes its blocks  (parallel)
//   5. tree_fill_kernel      the few tree blocks in planting order, minus those a LATER column's body overwrites
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "host/noise_impl.h"

namespace uvt {
namespace pg {

constexpr uint32_t kLcgA = 1103515245u, kLcgC = 12345u;
constexpr uint32_t kSolid = 0x10000000u;
constexpr int kMaxRanges = 4;            // tree-occupied height ranges remembered per column (more: the host path is used)
constexpr uint32_t kEvDeco = 4096u, kEvTree = 4097u;
constexpr unsigned long long kNoKey = ~0ull;

__host__ __device__ inline uint32_t lcg(uint32_t s) { return s * kLcgA + kLcgC; }

struct Tree {
    uint32_t x, z, y;   // column and base height (= vh of the column)
    uint32_t seed;      // LCG state before the trunk-height draw
};

__global__ void heights_kernel(const float *__restrict__ grad, uint32_t dim, float off_x, float off_y, uint16_t *__restrict__ vh) {
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
    if (z >= dim) return;
    vh[(size_t)x * dim + z] = (uint16_t)uvt_noise::column_height(grad, dim, x, z, off_x, off_y);
}

// n % d == 0 for odd d, without a division: n * d^-1 (mod 2^32) <= (2^32 - 1) / d
__host__ __device__ __forceinline__ bool divisible(uint32_t n, uint32_t inv, uint32_t limit) { return n * inv <= limit; }
constexpr uint32_t kInv5 = 0xCCCCCCCDu, kLim5 = 0xFFFFFFFFu / 5u;
constexpr uint32_t kInv71 = 0xE327A977u, kLim71 = 0xFFFFFFFFu / 71u;     // 71 * 0xE327A977 == 1 (mod 2^32)
constexpr uint32_t kInv105 = 0xD8FD8FD9u, kLim105 = 0xFFFFFFFFu / 105u;  // 420 = 4 * 105

// L^k(s) = A[k] * s + C[k], k = 0..5
struct LcgPow {
    uint32_t A[6], C[6];
    __host__ __device__ LcgPow() {
        A[0] = 1u; C[0] = 0u;
        for (int k = 1; k < 6; ++k) { A[k] = A[k - 1] * kLcgA; C[k] = C[k - 1] * kLcgA + kLcgC; }
    }
};

// body draws of a column: one per block, one more for sand (h <= 15) or for the top layer (procgen.zig:26-33)
__host__ __device__ __forceinline__ uint32_t body_draws(uint32_t h) { return h + (h < 16u ? h : 16u) + (h > 16u ? 1u : 0u); }

// ---- the LCG scan -------------------------------------------------------------------------------------------------
// One CTA, ONE lane carries the LCG state through the columns in the reference order; the other warps prepare, one tile
// ahead, what does not depend on that state.  For a column of height h, with (Aj, Cj) the jump over its body draws:
//   * h <= 16: no decoration, s' = L^(b+1)(s) (body + the trailing draw of procgen.zig:52);
//   * h > 16:  r1 = L(sb) decides (r1 % 5 == 0) between 5 and 4 further draws (grass type, r2, r3, trailing) — the test and
//     both outcomes are AFFINE in s with per-column constants: hit <=> M s + K <= (2^32 - 1) / 5, s' = Ah s + Ch or An s + Cn.
//     That is the whole dependent chain of an ordinary column: 3 multiply-adds, a compare, a select.
// Trees (r3 % 420 == 0) can only be planted for 5 < x, z < 500 and reach two columns further (procgen.zig:47, 55-70), so only
// tiles that intersect [6, 501]^2 take the general per-column code (tree test, the `continue` of procgen.zig:37-38 against the
// height ranges earlier trees occupy); W4 has 6 % of its columns there.
constexpr int kScanTile = 256, kScanThreads = 128;

struct ScanTile {
    uint4 hit[kScanTile];   // m, k (hit <=> m s + k <= kLim5), ah, ch (s' after a hit)
    uint4 miss[kScanTile];  // an, cn (s' after a miss; also r3 after a hit), a3, c3 (r3 after a miss)
    uint32_t h[kScanTile];
};

__global__ void __launch_bounds__(kScanThreads) scan_kernel(const uint16_t *__restrict__ vh, uint32_t dim, const uint2 *__restrict__ jump, uint32_t seed0,
                            uint32_t *__restrict__ seeds, uint8_t *__restrict__ skipped, Tree *__restrict__ trees, uint32_t max_trees,
                            uint32_t *__restrict__ n_trees_out, ushort2 *__restrict__ blocked, uint8_t *__restrict__ n_blocked,
                            uint32_t *__restrict__ status, uint32_t *__restrict__ final_seed) {
    __shared__ ScanTile tiles[2];
    const LcgPow P;
    const uint32_t tiles_per_row = (dim + kScanTile - 1) / kScanTile, n_tiles = tiles_per_row * dim;
    const uint32_t warp = threadIdx.x >> 5;

    auto prepare = [&](uint32_t t, ScanTile &T, uint32_t first, uint32_t stride) {
        const uint32_t x = t / tiles_per_row, z0 = (t % tiles_per_row) * kScanTile;
        for (uint32_t q = first; q < (uint32_t)kScanTile; q += stride) {
            if (z0 + q >= dim) {  // padding of the row's last tile: the identity
                T.hit[q] = make_uint4(0u, 0xFFFFFFFFu, 1u, 0u);
                T.miss[q] = make_uint4(1u, 0u, 1u, 0u);
                T.h[q] = 0u;
                continue;
            }
            const uint32_t h = vh[(size_t)x * dim + z0 + q];
            const uint32_t b = body_draws(h);
            T.h[q] = h;
            if (h <= 16u) {  // never a hit; s' = body + the trailing draw
                const uint2 j1 = __ldg(&jump[b + 1u]);
                T.hit[q] = make_uint4(0u, 0xFFFFFFFFu, j1.x, j1.y);
                T.miss[q] = make_uint4(j1.x, j1.y, 1u, 1u);  // (r3 = s + 1 is never tested: h <= 16 columns plant nothing)
            } else {
                const uint2 j = __ldg(&jump[b]);
                T.hit[q] = make_uint4(P.A[1] * j.x * kInv5, (P.A[1] * j.y + P.C[1]) * kInv5, P.A[5] * j.x, P.A[5] * j.y + P.C[5]);
                T.miss[q] = make_uint4(P.A[4] * j.x, P.A[4] * j.y + P.C[4], P.A[3] * j.x, P.A[3] * j.y + P.C[3]);
            }
        }
    };

    prepare(0, tiles[0], threadIdx.x, kScanThreads);
    __syncthreads();

    uint32_t s = seed0, n_trees = 0;
    bool row_dirty[3] = {false, false, false};             // n_blocked is zeroed by the caller
    uint32_t blk_lo[3] = {~0u, ~0u, ~0u}, blk_hi[3] = {0u, 0u, 0u};  // z range of the columns with tree ranges, per window row

    for (uint32_t t = 0; t < n_tiles; ++t) {
        if (warp != 0) {
            if (t + 1 < n_tiles) prepare(t + 1, tiles[(t + 1) & 1], threadIdx.x - 32u, kScanThreads - 32u);
        } else if (threadIdx.x == 0) {
            const ScanTile &T = tiles[t & 1];
            const uint32_t x = t / tiles_per_row, z0 = (t % tiles_per_row) * kScanTile;
            const uint32_t nz = min((uint32_t)kScanTile, dim - z0);
            uint32_t *out = seeds + (size_t)x * dim + z0;
            if (z0 == 0) {  // the row that enters the three-row window of tree ranges
                const uint32_t r = (x + 2u) % 3u;
                if (row_dirty[r]) {
                    for (uint32_t z = blk_lo[r]; z <= blk_hi[r]; ++z) n_blocked[r * dim + z] = 0;
                    row_dirty[r] = false;
                    blk_lo[r] = ~0u; blk_hi[r] = 0u;
                }
            }
            const bool general = x >= 6u && x <= 501u && z0 <= 501u && z0 + nz > 6u;
            if (!general) {
                // groups of four columns: the constants of the next group are fetched while this group's chain runs
                uint4 ha[4], mi[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { ha[u] = T.hit[u]; mi[u] = T.miss[u]; }
                for (uint32_t q = 0; q < nz; q += 4) {  // (rows are padded with identity columns up to the tile size; dim % 8 == 0)
                    uint4 hn[4], mn[4];
                    const uint32_t qn = (q + 4 < (uint32_t)kScanTile) ? q + 4 : q;
#pragma unroll
                    for (int u = 0; u < 4; ++u) { hn[u] = T.hit[qn + u]; mn[u] = T.miss[qn + u]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        out[q + u] = s;
                        const bool hit = ha[u].x * s + ha[u].y <= kLim5;
                        const uint32_t sh = ha[u].z * s + ha[u].w, sn = mi[u].x * s + mi[u].y;
                        s = hit ? sh : sn;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) { ha[u] = hn[u]; mi[u] = mn[u]; }
                }
            } else {
                const uint32_t r = x % 3u;
                uint32_t lo = blk_lo[r], hi = blk_hi[r];  // columns of this row that carry tree ranges (refreshed when a tree is planted)
                const bool plantable = x > 5u && x < 500u && x + 2u < dim;
                uint4 ha = T.hit[0], mi = T.miss[0];
                uint32_t h = T.h[0];
                for (uint32_t q = 0; q < nz; ++q) {
                    const uint32_t z = z0 + q;
                    const uint4 ha_n = T.hit[q + 1 < (uint32_t)kScanTile ? q + 1 : q], mi_n = T.miss[q + 1 < (uint32_t)kScanTile ? q + 1 : q];
                    const uint32_t h_n = T.h[q + 1 < (uint32_t)kScanTile ? q + 1 : q];
                    out[q] = s;
                    const bool hit = ha.x * s + ha.y <= kLim5;
                    const uint32_t sh = ha.z * s + ha.w, sn = mi.x * s + mi.y, s3 = mi.z * s + mi.w;
                    const uint32_t r3 = hit ? sn : s3;  // the % 420 draw: L^4 of the body end after a hit, L^3 after a miss
                    const uint32_t s_in = s;
                    s = hit ? sh : sn;
                    const bool tree = (r3 & 3u) == 0u && divisible(r3, kInv105, kLim105);
                    if (h > 16u && (tree || (z >= lo && z <= hi))) {  // rare: a tree to plant, or tree blocks over this column
                        bool blk = false;
                        if (z >= lo && z <= hi) {  // procgen.zig:37-38: a tree block already at (x, vh, z)?
                            for (uint32_t w = 0; w < n_blocked[r * dim + z]; ++w) {
                                const ushort2 rg = blocked[((size_t)r * dim + z) * kMaxRanges + w];
                                blk = blk || (h >= rg.x && h <= rg.y);
                            }
                        }
                        if (blk) {  // `continue`: the body draws only — no decoration draws, no trailing draw
                            const uint2 j = __ldg(&jump[body_draws(h)]);
                            skipped[(size_t)x * dim + z] = 1;
                            s = j.x * s_in + j.y;
                        } else if (tree && plantable && z > 5u && z < 500u && z + 2u < dim) {
                            // place_tree (procgen.zig:47-48, 55-70)
                            const uint32_t th = (r3 * kLcgA + kLcgC) % 4u + 4u;
                            if (n_trees < max_trees) trees[n_trees] = Tree{x, z, h, r3};
                            else *status = 2u;
                            ++n_trees;
                            for (uint32_t a = 0; a < 3; ++a)
                                for (uint32_t c = 0; c < 3; ++c) {
                                    if (a == 0 && c == 0) continue;  // this column is done
                                    const uint32_t ra = (x + a) % 3u;
                                    const size_t col = (size_t)ra * dim + (z + c);
                                    const uint32_t nb = n_blocked[col];
                                    if (nb >= (uint32_t)kMaxRanges) { *status = 1u; continue; }
                                    const uint32_t rlo = (a == 1 && c == 1) ? h : h + th;  // trunk column: trunk + canopy are one run
                                    blocked[col * kMaxRanges + nb] = make_ushort2((unsigned short)rlo, (unsigned short)(h + th + 2u));
                                    n_blocked[col] = (uint8_t)(nb + 1u);
                                    row_dirty[ra] = true;
                                    blk_lo[ra] = min(blk_lo[ra], z + c);
                                    blk_hi[ra] = max(blk_hi[ra], z + c);
                                }
                            lo = blk_lo[r]; hi = blk_hi[r];
                            const uint2 jt = __ldg(&jump[28u + th]);  // trunk height + th trunk types + 27 leaves
                            s = (jt.x * r3 + jt.y) * kLcgA + kLcgC;   // ... and the trailing draw
                        }
                    }
                    ha = ha_n; mi = mi_n; h = h_n;
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *n_trees_out = n_trees;
        *final_seed = s;
    }
}

// Decoration of every column from its start state (parallel): grass blade when r1 % 5 == 0, flower when r2 % 71 == 0
// (procgen.zig:40-45); none for columns the scan skipped (`continue`) or at most 16 high.
__global__ void deco_kernel(const uint16_t *__restrict__ vh, const uint32_t *__restrict__ seeds, const uint8_t *__restrict__ skipped,
                            const uint2 *__restrict__ jump, size_t n_col, uint32_t *__restrict__ deco) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_col) return;
    const uint32_t h = vh[i];
    uint32_t word = 0;
    if (h > 16u && !skipped[i]) {
        const uint2 j = __ldg(&jump[body_draws(h)]);
        uint32_t r = lcg(j.x * seeds[i] + j.y);  // r1
        if (divisible(r, kInv5, kLim5)) {
            r = lcg(r);
            word = 7u + r % 5u;                  // grass blade model, not solid
        }
        r = lcg(r);                              // r2
        if (divisible(r, kInv71, kLim71)) word = 12u | kSolid;
    }
    deco[i] = word;
}

__device__ __forceinline__ unsigned long long touch_key(uint32_t x, uint32_t z, uint32_t ev) {
    return ((unsigned long long)x << 40) | ((unsigned long long)z << 20) | ev;
}

// One thread per chunk column (cx, cz): the first body / decoration set() of every chunk above the water slab.
__global__ void touch_kernel(const uint16_t *__restrict__ vh, const uint32_t *__restrict__ deco, uint32_t dim, unsigned long long *__restrict__ keys) {
    const uint32_t cd = dim >> 3;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cd * cd) return;
    const uint32_t cx = t / cd, cz = t % cd;
    uint32_t top = 1;  // chunk rows 0 and 1 belong to the y < 16 slab (first loop of procgen.zig)
    for (uint32_t lx = 0; lx < 8; ++lx)
        for (uint32_t lz = 0; lz < 8; ++lz) {
            const uint32_t x = cx * 8 + lx, z = cz * 8 + lz;
            const size_t i = (size_t)x * dim + z;
            const uint32_t h = vh[i];
            if (h > 0) {
                const uint32_t cyb = min((h - 1) >> 3, cd - 1);
                while (top < cyb) {
                    ++top;
                    atomicMin(&keys[cx + (size_t)cd * (top + (size_t)cz * cd)], touch_key(x, z, top * 8));  // block h = 8 * top of this column
                }
            }
            if (deco[i] != 0 && (h >> 3) < cd && (h >> 3) > top) {
                top = h >> 3;
                atomicMin(&keys[cx + (size_t)cd * (top + (size_t)cz * cd)], touch_key(x, z, kEvDeco));
            }
        }
}

// The blocks of one tree in set() order (procgen.zig:55-70): k = 0 trunk base, 1..th trunk, then the 27 leaves (a, b, c nested).
struct TreeBlock {
    uint32_t x, y, z, word;
};
__device__ __forceinline__ uint32_t tree_blocks(const Tree &t, TreeBlock out[36]) {
    uint32_t s = lcg(t.seed);
    const uint32_t th = s % 4u + 4u;
    uint32_t n = 0;
    out[n++] = TreeBlock{t.x + 1, t.y, t.z + 1, 15u | kSolid};
    for (uint32_t off = 0; off < th; ++off) {
        s = lcg(s);
        out[n++] = TreeBlock{t.x + 1, t.y + off, t.z + 1, (14u + s % 3u) | kSolid};
    }
    for (uint32_t a = 0; a < 3; ++a)
        for (uint32_t b = 0; b < 3; ++b)
            for (uint32_t c = 0; c < 3; ++c) {
                s = lcg(s);
                out[n++] = TreeBlock{t.x + a, t.y + th + b, t.z + c, (18u + s % 2u) | kSolid};
            }
    return n;
}

__global__ void tree_touch_kernel(const Tree *__restrict__ trees, uint32_t n_trees, uint32_t dim, unsigned long long *__restrict__ keys,
                                  uint32_t *__restrict__ status) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    const uint32_t cd = dim >> 3;
    TreeBlock blk[36];
    const uint32_t n = tree_blocks(trees[t], blk);
    for (uint32_t k = 0; k < n; ++k) {
        if (blk[k].x >= dim || blk[k].y >= dim || blk[k].z >= dim) { *status = 3u; continue; }  // the host set() would fail: host path reports it
        atomicMin(&keys[(blk[k].x >> 3) + (size_t)cd * ((blk[k].y >> 3) + (size_t)(blk[k].z >> 3) * cd)], touch_key(trees[t].x, trees[t].z, kEvTree + k));
    }
}

__global__ void iota_kernel(uint32_t *v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// chunk table: the slab bricks in (cx, cz, cy) order, then the sorted first-touch order
__global__ void slab_chunks_kernel(uint32_t *__restrict__ chunks, uint32_t cd) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cd * cd * 2u) return;
    const uint32_t cy = t & 1u, cz = (t >> 1) % cd, cx = (t >> 1) / cd;
    chunks[cx + cd * (cy + cz * cd)] = t + 1u;
}
__global__ void ranked_chunks_kernel(const unsigned long long *__restrict__ sorted_keys, const uint32_t *__restrict__ sorted_chunk, uint32_t n,
                                     uint32_t first_brick, uint32_t *__restrict__ chunks, uint32_t *__restrict__ n_touched) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || sorted_keys[i] == kNoKey) return;
    chunks[sorted_chunk[i]] = first_brick + i + 1u;
    atomicMax(n_touched, i + 1u);
}

// One thread per column; x is the fast thread index so that the 8 columns of a brick row store one 32-byte run.
__global__ void fill_kernel(const uint16_t *__restrict__ vh, const uint32_t *__restrict__ seeds, const uint32_t *__restrict__ deco,
                            const uint32_t *__restrict__ chunks, uint32_t dim, uint32_t *__restrict__ bricks) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y;
    if (x >= dim) return;
    const uint32_t cd = dim >> 3;
    const size_t i = (size_t)x * dim + z;
    const uint32_t h_col = vh[i];
    uint32_t s = seeds[i];
    const uint32_t in_brick = (x & 7u) + 64u * (z & 7u);
    const size_t chunk_col = (x >> 3) + (size_t)cd * cd * (z >> 3);
    uint32_t *brick = nullptr;
    const uint32_t top = max(h_col, 16u);
    for (uint32_t y = 0; y < top && y < dim; ++y) {
        if ((y & 7u) == 0u) brick = bricks + (size_t)(chunks[chunk_col + (size_t)cd * (y >> 3)] - 1u) * 512u;
        uint32_t word = 13u | kSolid;                                  // water slab (procgen.zig:10-19) where the terrain is lower
        if (y < h_col) {
            s = lcg(s);
            word = (21u + s % 3u) | kSolid;                            // procgen.zig:27
            if (y <= 15u) { s = lcg(s); word = (25u + s % 3u) | kSolid; }                         // :29-30 sand
            else if (y == h_col - 1u) { s = lcg(s); word = (s % 6u) | kSolid; }                   // :31-32 top layer
        }
        brick[in_brick + 8u * (y & 7u)] = word;
    }
    const uint32_t d = deco[i];
    if (d != 0 && h_col < dim) bricks[(size_t)(chunks[chunk_col + (size_t)cd * (h_col >> 3)] - 1u) * 512u + in_brick + 8u * (h_col & 7u)] = d;
}

// ONE thread: the tree blocks in planting order.  A block over a column that comes LATER in the reference order is overwritten
// by that column's body when it lies under its height (the body runs after the tree was planted).
__global__ void tree_fill_kernel(const Tree *__restrict__ trees, uint32_t n_trees, const uint16_t *__restrict__ vh, const uint32_t *__restrict__ chunks,
                                 uint32_t dim, uint32_t *__restrict__ bricks) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const uint32_t cd = dim >> 3;
    TreeBlock blk[36];
    for (uint32_t t = 0; t < n_trees; ++t) {
        const Tree tr = trees[t];
        const uint32_t n = tree_blocks(tr, blk);
        for (uint32_t k = 0; k < n; ++k) {
            const TreeBlock b = blk[k];
            if (b.x >= dim || b.y >= dim || b.z >= dim) continue;
            const bool later = b.x > tr.x || (b.x == tr.x && b.z > tr.z);
            if (later && b.y < vh[(size_t)b.x * dim + b.z]) continue;
            const uint32_t e = chunks[(b.x >> 3) + (size_t)cd * ((b.y >> 3) + (size_t)(b.z >> 3) * cd)];
            bricks[(size_t)(e - 1u) * 512u + (b.x & 7u) + 8u * (b.y & 7u) + 64u * (b.z & 7u)] = b.word;
        }
    }
}

}  // namespace pg
}  // namespace uvt
