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
__device__ __forceinline__ bool divisible(uint32_t n, uint32_t inv, uint32_t limit) { return n * inv <= limit; }

// ONE thread.  jump[n] = (A_n, C_n) with L^n(s) = A_n * s + C_n.  blocked[3][dim][kMaxRanges] / n_blocked[3][dim]: height ranges
// of tree blocks over the columns of rows x, x + 1, x + 2 (a tree reaches two columns ahead).  status: 1 = range overflow.
__global__ void scan_kernel(const uint16_t *__restrict__ vh, uint32_t dim, const uint2 *__restrict__ jump, uint32_t seed0,
                            uint32_t *__restrict__ seeds, uint32_t *__restrict__ deco, Tree *__restrict__ trees, uint32_t max_trees,
                            uint32_t *__restrict__ n_trees_out, ushort2 *__restrict__ blocked, uint8_t *__restrict__ n_blocked,
                            uint32_t *__restrict__ status, uint32_t *__restrict__ final_seed) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    constexpr uint32_t inv5 = 0xCCCCCCCDu, lim5 = 0xFFFFFFFFu / 5u;
    constexpr uint32_t inv71 = 0xE327A977u, lim71 = 0xFFFFFFFFu / 71u;     // 71 * 0xE327A977 == 1 (mod 2^32)
    constexpr uint32_t inv105 = 0xD8FD8FD9u, lim105 = 0xFFFFFFFFu / 105u;  // 420 = 4 * 105
    // affine powers of the LCG step
    uint32_t A[6], C[6];
    A[0] = 1u; C[0] = 0u;
    for (int k = 1; k < 6; ++k) { A[k] = A[k - 1] * kLcgA; C[k] = C[k - 1] * kLcgA + kLcgC; }
    uint32_t s = seed0, n_trees = 0;
    int last_tree_x = -100;
    bool row_dirty[3] = {false, false, false};  // n_blocked is zeroed by the caller
    for (uint32_t x = 0; x < dim; ++x) {
        {   // the row that enters the three-row window
            const uint32_t r = (x + 2u) % 3u;
            if (row_dirty[r]) {
                for (uint32_t z = 0; z < dim; ++z) n_blocked[r * dim + z] = 0;
                row_dirty[r] = false;
            }
        }
        const bool near_tree = last_tree_x + 2 >= (int)x;
        const uint16_t *row = vh + (size_t)x * dim;
        for (uint32_t z0 = 0; z0 < dim; z0 += 4) {
            // heights and jump entries of four columns first: their latency is off the LCG chain
            const uint2 hh = *reinterpret_cast<const uint2 *>(row + z0);
            const uint32_t h4[4] = {hh.x & 0xFFFFu, hh.x >> 16, hh.y & 0xFFFFu, hh.y >> 16};
            uint2 j4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t h = h4[k];
                j4[k] = __ldg(&jump[h + min(h, 16u) + (h > 16u ? 1u : 0u)]);  // body draws: one per block, one more for sand (h <= 15) / the top layer
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t z = z0 + k, h = h4[k];
                const size_t i = (size_t)x * dim + z;
                seeds[i] = s;
                const uint32_t sb = j4[k].x * s + j4[k].y;  // state after the column body
                if (h <= 16u) {                             // no decoration: the trailing draw only (procgen.zig:52)
                    s = sb * kLcgA + kLcgC;
                    continue;
                }
                if (near_tree || last_tree_x + 2 >= (int)x) {  // procgen.zig:37-38: something (a tree block) already at (x, vh, z)
                    const uint32_t r = x % 3u;
                    bool blk = false;
                    for (uint32_t q = 0; q < n_blocked[r * dim + z]; ++q) {
                        const ushort2 rg = blocked[((size_t)r * dim + z) * kMaxRanges + q];
                        blk = blk || (h >= rg.x && h <= rg.y);
                    }
                    if (blk) {  // `continue`: no decoration draws and no trailing draw
                        s = sb;
                        continue;
                    }
                }
                // draws after the body: r1 (% 5), [grass type], r2 (% 71), r3 (% 420), [tree], trailing — every candidate is an
                // affine function of sb, so they are all evaluated side by side and selected by the % 5 outcome
                const uint32_t r1 = A[1] * sb + C[1];
                const bool hit5 = divisible(r1, inv5, lim5);
                const uint32_t g = A[2] * sb + C[2];                 // grass type draw when hit5, else r2
                const uint32_t r2 = hit5 ? A[3] * sb + C[3] : g;
                const uint32_t r3a = A[3] * sb + C[3], r3b = A[4] * sb + C[4];
                const uint32_t r3 = hit5 ? r3b : r3a;
                const bool t420 = hit5 ? ((r3b & 3u) == 0u && divisible(r3b >> 2, inv105, lim105)) : ((r3a & 3u) == 0u && divisible(r3a >> 2, inv105, lim105));
                uint32_t word = 0;
                if (hit5) word = 7u + g % 5u;                            // grass blade model, not solid (procgen.zig:42)
                if (divisible(r2, inv71, lim71)) word = 12u | kSolid;    // flower (procgen.zig:45)
                if (word) deco[i] = word;
                s = (hit5 ? A[5] * sb + C[5] : A[4] * sb + C[4]);        // ... and the trailing draw
                if (t420 && x < 500u && z < 500u && x > 5u && z > 5u && x + 2u < dim && z + 2u < dim) {  // place_tree (procgen.zig:47-48, 55-70)
                    const uint32_t th = (r3 * kLcgA + kLcgC) % 4u + 4u;
                    if (n_trees < max_trees) trees[n_trees] = Tree{x, z, h, r3};
                    else *status = 2u;
                    ++n_trees;
                    last_tree_x = (int)x;
                    for (uint32_t a = 0; a < 3; ++a)
                        for (uint32_t c = 0; c < 3; ++c) {
                            if (a == 0 && c == 0) continue;  // this column is done
                            const uint32_t r = (x + a) % 3u;
                            const size_t col = (size_t)r * dim + (z + c);
                            const uint32_t nb = n_blocked[col];
                            if (nb >= (uint32_t)kMaxRanges) { *status = 1u; continue; }
                            const uint32_t lo = (a == 1 && c == 1) ? h : h + th;  // trunk column: trunk + canopy are one run
                            blocked[col * kMaxRanges + nb] = make_ushort2((unsigned short)lo, (unsigned short)(h + th + 2u));
                            n_blocked[col] = (uint8_t)(nb + 1u);
                            row_dirty[r] = true;
                        }
                    const uint2 jt = __ldg(&jump[28u + th]);  // trunk height + th trunk types + 27 leaves
                    s = (jt.x * r3 + jt.y) * kLcgA + kLcgC;   // ... and the trailing draw
                }
            }
        }
    }
    *n_trees_out = n_trees;
    *final_seed = s;
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
