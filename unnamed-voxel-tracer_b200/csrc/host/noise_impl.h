// noise_impl.h — FastNoiseLite-style OpenSimplex2 FBm, the height source of the reference procgen
// (src/procgen.zig:7,23: znoise.FnlGenerator{ .fractal_type = .fbm }.noise2), written once for the host
// (csrc/host/noise.cpp) and for the device (procgen kernels in csrc/uvt.cu): every operation is an individually
// rounded IEEE fp32 op on both sides (gcc -ffp-contract=off, nvcc --fmad=false), so the two agree bit for bit.
//
// znoise @96f9458c (build.zig.zon:19-22) wraps the FastNoiseLite C library, which is NOT vendored under the
// reference tree.  This restates the library's published algorithm with its documented defaults (seed 1337,
// frequency 0.01, OpenSimplex2, 3 octaves, lacunarity 2, gain 0.5, weighted strength 0).  PARITY UNPINNED: there
// is no copy of the library here to compare bytes against; the world is an INPUT of the traversal path, and the
// same generated world feeds the oracle and the kernels.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define UVT_HD __host__ __device__
#else
#define UVT_HD
#endif

namespace uvt_noise {

constexpr int kSeed = 1337;
constexpr float kFrequency = 0.01f;
constexpr int kOctaves = 3;
constexpr float kLacunarity = 2.0f;
constexpr float kGain = 0.5f;
constexpr int kPrimeX = 501125321;
constexpr int kPrimeY = 1136930381;

// 24 unit vectors at 7.5 deg + 15 deg * k (clockwise from +y), tiled 5x, then 8 diagonal fillers at
// 22.5 deg + 45 deg * k: 128 gradients (256 floats), built on the host from the 6 base magnitudes.
inline void build_grad_table(float g[256]) {
    static const float m[6] = {0.130526192220052f, 0.38268343236509f, 0.608761429008721f,
                               0.793353340291235f, 0.923879532511287f, 0.99144486137381f};
    float ring[48];
    for (int k = 0; k < 24; ++k) {
        int q = k / 6, r = k % 6;   // quadrant, step inside the quadrant
        float s = m[r], c = m[5 - r];
        float x, y;
        switch (q) {
            case 0: x = s; y = c; break;    // 7.5..82.5 deg from +y toward +x
            case 1: x = c; y = -s; break;
            case 2: x = -s; y = -c; break;
            default: x = -c; y = s; break;
        }
        ring[2 * k] = x;
        ring[2 * k + 1] = y;
    }
    for (int rep = 0; rep < 5; ++rep)
        for (int i = 0; i < 48; ++i) g[rep * 48 + i] = ring[i];
    const float d[16] = {m[1], m[4], m[4], m[1], m[4], -m[1], m[1], -m[4],
                         -m[1], -m[4], -m[4], -m[1], -m[4], m[1], -m[1], m[4]};
    for (int i = 0; i < 16; ++i) g[240 + i] = d[i];
}

UVT_HD inline int fast_floor(float f) { return f >= 0 ? (int)f : (int)f - 1; }

UVT_HD inline float grad_coord(const float *table, int seed, int xp, int yp, float xd, float yd) {
    // wrapping 32-bit hash
    uint32_t h = (uint32_t)seed ^ (uint32_t)xp ^ (uint32_t)yp;
    h *= 0x27d4eb2du;
    int hash = (int)h;
    hash ^= hash >> 15;
    hash &= 127 << 1;
    return xd * table[hash] + yd * table[hash | 1];
}

UVT_HD inline float single_simplex2(const float *table, int seed, float x, float y) {
    const float SQRT3 = 1.7320508075688772935274463415059f;
    const float G2 = (3 - SQRT3) / 6;

    int i = fast_floor(x);
    int j = fast_floor(y);
    float xi = (float)(x - i);
    float yi = (float)(y - j);

    float t = (xi + yi) * G2;
    float x0 = (float)(xi - t);
    float y0 = (float)(yi - t);

    i = (int)((uint32_t)i * (uint32_t)kPrimeX);
    j = (int)((uint32_t)j * (uint32_t)kPrimeY);

    float n0, n1, n2;

    float a = 0.5f - x0 * x0 - y0 * y0;
    if (a <= 0) n0 = 0;
    else n0 = (a * a) * (a * a) * grad_coord(table, seed, i, j, x0, y0);

    float c = (float)(2 * (1 - 2 * G2) * (1 / G2 - 2)) * t + ((float)(-2 * (1 - 2 * G2) * (1 - 2 * G2)) + a);
    if (c <= 0) n2 = 0;
    else {
        float x2 = x0 + (2 * (float)G2 - 1);
        float y2 = y0 + (2 * (float)G2 - 1);
        n2 = (c * c) * (c * c) * grad_coord(table, seed, (int)((uint32_t)i + (uint32_t)kPrimeX), (int)((uint32_t)j + (uint32_t)kPrimeY), x2, y2);
    }

    if (y0 > x0) {
        float x1 = x0 + (float)G2;
        float y1 = y0 + ((float)G2 - 1);
        float b = 0.5f - x1 * x1 - y1 * y1;
        if (b <= 0) n1 = 0;
        else n1 = (b * b) * (b * b) * grad_coord(table, seed, i, (int)((uint32_t)j + (uint32_t)kPrimeY), x1, y1);
    } else {
        float x1 = x0 + ((float)G2 - 1);
        float y1 = y0 + (float)G2;
        float b = 0.5f - x1 * x1 - y1 * y1;
        if (b <= 0) n1 = 0;
        else n1 = (b * b) * (b * b) * grad_coord(table, seed, (int)((uint32_t)i + (uint32_t)kPrimeX), j, x1, y1);
    }

    return (n0 + n1 + n2) * 99.83685446303647f;
}

UVT_HD inline float fractal_bounding() {
    float gain = kGain < 0 ? -kGain : kGain;
    float amp = gain;
    float amp_fractal = 1.0f;
    for (int i = 1; i < kOctaves; ++i) {
        amp_fractal += amp;
        amp *= gain;
    }
    return 1.0f / amp_fractal;
}

UVT_HD inline float noise2_fbm(const float *table, float x, float y) {
    // coordinate transform: frequency, then the OpenSimplex2 skew
    x *= kFrequency;
    y *= kFrequency;
    {
        const float SQRT3 = 1.7320508075688772935274463415059f;
        const float F2 = 0.5f * (SQRT3 - 1);
        float t = (x + y) * F2;
        x += t;
        y += t;
    }
    int seed = kSeed;
    float sum = 0;
    float amp = fractal_bounding();
    for (int i = 0; i < kOctaves; ++i) {
        float noise = single_simplex2(table, seed++, x, y);
        sum += noise * amp;
        // weighted strength 0: amp *= lerp(1, min(noise+1,2)*0.5, 0) == amp * 1
        x *= kLacunarity;
        y *= kLacunarity;
        amp *= kGain;
    }
    return sum;
}

// procgen.zig:23-24: noise2((offX + x)/10, (offY + z)/10); vh = u32(max(val * dim * 0.1, 0))
UVT_HD inline uint32_t column_height(const float *table, uint32_t dim, uint32_t x, uint32_t z, float offset_x, float offset_y) {
    const float val = noise2_fbm(table, (offset_x + (float)x) / 10.0f, (offset_y + (float)z) / 10.0f);
    float h = val * (float)dim * 0.1f;
    if (!(h > 0.0f)) h = 0.0f;
    return (uint32_t)h;
}

}  // namespace uvt_noise
