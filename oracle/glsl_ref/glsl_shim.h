// glsl_shim.h — just enough of GLSL 4.50 in C++20 to compile the reference's OWN shader text on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see oracle/glsl_ref/README.md).  oracle/glsl_ref/translate.py reads the shaders where they
// lie under /root/reference/assets/shaders, expands their #include lines like the reference loader does
// (src/engine/graphics/shader.zig:14-40) and applies a handful of purely syntactic rewrites (listed in translate.py);
// the result is compiled against this header into oracle/_ref/libglslref.so.  What the shim supplies is the GLSL
// *library*: vector types, operators and built-ins.  Its arithmetic follows the contract SURVEY App. A fixes for
// everything GLSL leaves implementation-defined: fp32 throughout with every operation individually rounded (build with
// -ffp-contract=off, no fast-math), float->int conversions truncate toward zero, saturate and send NaN to 0, min/max
// are the specification's ternaries, mat4 * vec4 and dot() sum left to right, normalize() divides by sqrt(dot),
// RGBA8 image stores clamp, scale by 255, add 0.5 and truncate.  One choice differs from oracle.c on purpose: clamp()
// is NaN-suppressing like the min/max of real GPUs, which is what keeps the blit's `normalize(vec3(0))` on sky pixels
// from blackening the sky (oracle.c states the same outcome as a convention).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef unsigned int uint;

inline int f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int)f;
}
inline uint f2u(float f) {
    if (f != f || f <= 0.0f) return 0u;
    if (f >= 4294967296.0f) return UINT32_MAX;
    return (uint)f;
}

struct ivec2;
struct ivec3;
struct uvec2;
struct uvec3;
struct vec3;

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(const ivec2 &v);
    vec2 xy() const { return *this; }
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    explicit vec3(int s) : x((float)s), y((float)s), z((float)s) {}
    explicit vec3(double s) : x((float)s), y((float)s), z((float)s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(int a, int b, int c) : x((float)a), y((float)b), z((float)c) {}
    explicit vec3(const ivec3 &v);
    vec3 xyz() const { return *this; }
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

struct vec4 {
    union { float x; float r; };
    union { float y; float g; };
    union { float z; float b; };
    union { float w; float a; };
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
    vec4(const vec3 &v, float d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    vec4(const vec2 &v, float c_, float d_) : x(v.x), y(v.y), z(c_), w(d_) {}
    vec3 xyz() const { return vec3(x, y, z); }
    vec2 xy() const { return vec2(x, y); }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const uvec2 &v);
};
struct uvec2 {
    uint x, y;
    uvec2(uint a, uint b) : x(a), y(b) {}
};

struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    explicit ivec3(int s) : x(s), y(s), z(s) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    ivec3(uint a, uint b, uint c) : x((int)a), y((int)b), z((int)c) {}
    explicit ivec3(const vec3 &v) : x(f2i(v.x)), y(f2i(v.y)), z(f2i(v.z)) {}
    explicit ivec3(const uvec3 &v);
    int &operator[](int i) { return (&x)[i]; }
    int operator[](int i) const { return (&x)[i]; }
};

struct uvec3 {
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    explicit uvec3(const ivec3 &v) : x((uint)v.x), y((uint)v.y), z((uint)v.z) {}
    explicit uvec3(const vec3 &v) : x(f2u(v.x)), y(f2u(v.y)), z(f2u(v.z)) {}
    uvec2 xy() const { return uvec2(x, y); }
};

struct bvec3 {
    bool x, y, z;
};

inline vec2::vec2(const ivec2 &v) : x((float)v.x), y((float)v.y) {}
inline vec3::vec3(const ivec3 &v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
inline ivec2::ivec2(const uvec2 &v) : x((int)v.x), y((int)v.y) {}
inline ivec3::ivec3(const uvec3 &v) : x((int)v.x), y((int)v.y), z((int)v.z) {}

// ---- scalar built-ins (GLSL 4.50 §8; min/max per the specification's formulas) --------------------
inline float min(float x, float y) { return y < x ? y : x; }
inline float max(float x, float y) { return x < y ? y : x; }
inline float max(float x, int y) { return max(x, (float)y); }
inline float min(float x, int y) { return min(x, (float)y); }
inline int min(int x, int y) { return y < x ? y : x; }
inline int max(int x, int y) { return x < y ? y : x; }
inline float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }  // NaN-suppressing, see the header note
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float floor(float x) { return floorf(x); }
inline float ceil(float x) { return ceilf(x); }
inline float fract(float x) { return x - floorf(x); }
inline float sqrt(float x) { return sqrtf(x); }
inline float pow(float x, float y) { return powf(x, y); }
inline float tan(float x) { return tanf(x); }
inline float sin(float x) { return sinf(x); }
inline float cos(float x) { return cosf(x); }
inline float abs(float x) { return fabsf(x); }
inline float mod(float x, float y) { return x - y * floorf(x / y); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }

// ---- vec2 ----------------------------------------------------------------------------------------------
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator*(int s, vec2 a) { return vec2((float)s * a.x, (float)s * a.y); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(float s, vec2 a) { return vec2(s - a.x, s - a.y); }
inline vec2 &operator*=(vec2 &a, float s) { a = a * s; return a; }
inline vec2 &operator*=(vec2 &a, vec2 b) { a = a * b; return a; }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float length(vec2 a) { return sqrtf(dot(a, a)); }

// ---- vec3 ----------------------------------------------------------------------------------------------
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, vec3 a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 &operator+=(vec3 &a, vec3 b) { a = a + b; return a; }
inline vec3 &operator*=(vec3 &a, float s) { a = a * s; return a; }
// int -> float is GLSL's one implicit conversion: ivec3 operands promote to vec3
inline vec3 operator+(ivec3 a, vec3 b) { return vec3(a) + b; }
inline vec3 operator+(vec3 a, ivec3 b) { return a + vec3(b); }
inline vec3 operator-(ivec3 a, vec3 b) { return vec3(a) - b; }
inline vec3 operator-(vec3 a, ivec3 b) { return a - vec3(b); }
inline vec3 &operator+=(vec3 &a, ivec3 b) { a = a + vec3(b); return a; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline vec3 min(vec3 a, vec3 b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(vec3 a, vec3 b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 sign(vec3 a) { return vec3(sign(a.x), sign(a.y), sign(a.z)); }
inline vec3 fract(vec3 a) { return vec3(fract(a.x), fract(a.y), fract(a.z)); }
inline vec3 floor(vec3 a) { return vec3(floorf(a.x), floorf(a.y), floorf(a.z)); }
inline vec3 ceil(vec3 a) { return vec3(ceilf(a.x), ceilf(a.y), ceilf(a.z)); }
inline bvec3 lessThan(vec3 a, vec3 b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }

// ---- vec4 ----------------------------------------------------------------------------------------------
inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(vec4 a, vec4 b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(vec4 a, vec4 b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator*(float s, vec4 a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline vec4 operator/(vec4 a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline vec4 operator+(vec4 a, float s) { return vec4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline vec4 operator-(vec4 a, float s) { return vec4(a.x - s, a.y - s, a.z - s, a.w - s); }
inline vec4 &operator+=(vec4 &a, vec4 b) { a = a + b; return a; }
inline float dot(vec4 a, vec4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(vec4 a) { return sqrtf(dot(a, a)); }
inline vec4 normalize(vec4 a) { return a / length(a); }
inline vec4 fract(vec4 a) { return vec4(fract(a.x), fract(a.y), fract(a.z), fract(a.w)); }
inline vec4 mix(vec4 x, vec4 y, float a) { return x * (1.0f - a) + y * a; }

struct mat4 {
    float c[4][4];  // c[j] = column j
};
inline vec4 operator*(const mat4 &m, vec4 v) {
    vec4 r;
    for (int i = 0; i < 4; ++i) {
        float acc = m.c[0][i] * v.x;
        acc = acc + m.c[1][i] * v.y;
        acc = acc + m.c[2][i] * v.z;
        acc = acc + m.c[3][i] * v.w;
        r[i] = acc;
    }
    return r;
}

// ---- integer vectors -----------------------------------------------------------------------------------
inline ivec3 operator+(ivec3 a, ivec3 b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator-(ivec3 a, ivec3 b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ivec3 operator+(int s, ivec3 a) { return ivec3(s + a.x, s + a.y, s + a.z); }
inline ivec3 operator-(int s, ivec3 a) { return ivec3(s - a.x, s - a.y, s - a.z); }
inline ivec3 operator>>(ivec3 a, int s) { return ivec3(a.x >> s, a.y >> s, a.z >> s); }
inline ivec3 operator<<(ivec3 a, int s) { return ivec3(a.x << s, a.y << s, a.z << s); }
inline ivec3 operator<<(ivec3 a, uint s) { return ivec3(a.x << s, a.y << s, a.z << s); }
inline ivec3 operator&(ivec3 a, int m) { return ivec3(a.x & m, a.y & m, a.z & m); }
inline ivec3 operator%(ivec3 a, int m) { return ivec3(a.x % m, a.y % m, a.z % m); }
inline ivec3 &operator+=(ivec3 &a, ivec3 b) { a = a + b; return a; }
inline ivec3 &operator-=(ivec3 &a, ivec3 b) { a = a - b; return a; }
inline uvec3 operator+(uvec3 a, uvec3 b) { return uvec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline bvec3 lessThan(ivec3 a, ivec3 b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bvec3 greaterThanEqual(ivec3 a, ivec3 b) { return bvec3{a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline bool any(bvec3 v) { return v.x || v.y || v.z; }

// ---- packing (GLSL 4.50 §8.4) ----------------------------------------------------------------------------
inline uint packUnorm4x8(vec4 v) {
    uint out = 0;
    for (int i = 0; i < 4; ++i) out |= (uint)roundf(fminf(fmaxf(v[i], 0.0f), 1.0f) * 255.0f) << (8 * i);
    return out;
}
inline vec4 unpackUnorm4x8(uint p) {
    return vec4((float)(p & 255u) / 255.0f, (float)((p >> 8) & 255u) / 255.0f, (float)((p >> 16) & 255u) / 255.0f, (float)(p >> 24) / 255.0f);
}

// ---- images ------------------------------------------------------------------------------------------------
// RGBA8 UNORM texel store: clamp, *255, +0.5, truncate (the float->UNORM rule GPUs implement)
inline uint store_unorm8(float c) {
    if (c != c) return 0u;
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint)(c * 255.0f + 0.5f);
}
enum { FMT_RGBA8 = 0, FMT_RGBA32F = 1 };
struct image2D {
    int w = 0, h = 0, fmt = FMT_RGBA8;
    void *data = nullptr;  // row 0 = bottom row (GL image origin)
};
typedef image2D sampler2D;
struct image3D {
    int w = 0, h = 0, d = 0;
    const uint *data = nullptr;  // RGBA8, x fastest
};
inline ivec2 imageSize(const image2D &im) { return ivec2(im.w, im.h); }
inline vec4 imageLoad(const image2D &im, ivec2 p) {
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return vec4(0.0f);
    const size_t i = (size_t)p.y * im.w + p.x;
    if (im.fmt == FMT_RGBA32F) {
        const float *f = (const float *)im.data + 4 * i;
        return vec4(f[0], f[1], f[2], f[3]);
    }
    return unpackUnorm4x8(((const uint *)im.data)[i]);
}
inline void imageStore(image2D &im, ivec2 p, vec4 v) {
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return;
    const size_t i = (size_t)p.y * im.w + p.x;
    if (im.fmt == FMT_RGBA32F) {
        float *f = (float *)im.data + 4 * i;
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
        ((uint *)im.data)[i] = store_unorm8(v.x) | (store_unorm8(v.y) << 8) | (store_unorm8(v.z) << 16) | (store_unorm8(v.w) << 24);
    }
}
inline vec4 imageLoad(const image3D &im, ivec3 p) {
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= im.w || p.y >= im.h || p.z >= im.d) return vec4(0.0f);  // out-of-bounds image loads return 0
    return unpackUnorm4x8(im.data[(size_t)p.x + (size_t)im.w * ((size_t)p.y + (size_t)im.h * (size_t)p.z)]);
}
// texture(): the blit samples at fragment centres of a target as large as the textures, so every filter returns the texel
inline vec4 texture(const sampler2D &s, vec2 uv) {
    int x = (int)floorf(uv.x * (float)s.w), y = (int)floorf(uv.y * (float)s.h);
    x = x < 0 ? 0 : (x >= s.w ? s.w - 1 : x);
    y = y < 0 ? 0 : (y >= s.h ? s.h - 1 : y);
    return imageLoad(s, ivec2(x, y));
}

}  // namespace glsl
