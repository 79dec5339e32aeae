// glsl.hpp -- just enough GLSL 4.50 semantics, written from the GLSL specification,
// to compile the reference's compute-shader TEXT (resources/*.comp, *.glsl, read where
// it lies under /root/reference and lightly rewritten by glsl2cpp.py) as C++.
//
// TEST INFRASTRUCTURE ONLY.  This is how the CPU oracle (oracle/xn_oracle.c) is pinned
// against the reference's own source: oracle/_ref/libxnref_glsl.so runs the shader
// text itself; tests compare its images with the oracle's.
//
// Semantics chosen where GLSL leaves latitude (and mirrored by xn_oracle.c):
//   * all float arithmetic is IEEE binary32, no contraction (-ffp-contract=off);
//   * normalize(v) = v / sqrt(dot(v, v)); dot is evaluated left to right;
//   * imageStore to rgba8: clamp to [0,1], * 255, round half to even; NaN -> 0;
//   * texelFetch outside the image returns 0 (border / robust-access behaviour);
//   * shifts by >= 32 yield 0 (undefined in GLSL; only reachable on the ESVO
//     POP-underflow path whose result the loop condition discards).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

using uint = uint32_t;

struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };

struct vec2; struct ivec2; struct uvec2;
struct vec3; struct ivec3; struct uvec3;

struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int x_, int y_) : x(x_), y(y_) {}
    explicit ivec2(const uvec2& v);
};
struct uvec2 {
    uint x, y;
    uvec2() : x(0), y(0) {}
    uvec2(uint x_, uint y_) : x(x_), y(y_) {}
};
inline ivec2::ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}
struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    vec2(float x_, float y_) : x(x_), y(y_) {}
    explicit vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
    explicit vec2(const uvec2& v) : x((float)v.x), y((float)v.y) {}
    vec2& operator-=(float s) { x -= s; y -= s; return *this; }
};
inline ivec2 operator+(ivec2 a, ivec2 b) { return {a.x + b.x, a.y + b.y}; }
inline ivec2 operator-(ivec2 a, ivec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator/(vec2 a, vec2 b) { return {a.x / b.x, a.y / b.y}; }

struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    explicit ivec3(int s) : x(s), y(s), z(s) {}
    ivec3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
    explicit ivec3(const vec3& v); // truncation toward zero
    ivec3& operator+=(ivec3 b) { x += b.x; y += b.y; z += b.z; return *this; }
};
struct uvec3 {
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    explicit uvec3(uint s) : x(s), y(s), z(s) {}
    uvec3(uint x_, uint y_, uint z_) : x(x_), y(y_), z(z_) {}
    explicit uvec3(const vec3& v);
    explicit uvec3(const bvec3& b) : x(b.x), y(b.y), z(b.z) {}
    uvec2 xy() const { return {x, y}; }
    uvec3& operator%=(uint m) { x %= m; y %= m; z %= m; return *this; }
};
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    vec3(const ivec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {} // implicit in GLSL
    vec3(const uvec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {} // implicit in GLSL
    explicit vec3(const bvec3& b) : x(b.x ? 1.f : 0.f), y(b.y ? 1.f : 0.f), z(b.z ? 1.f : 0.f) {}
    vec3 xyz() const { return *this; }
    vec3 yzx() const { return {y, z, x}; }
    vec3 zxy() const { return {z, x, y}; }
    vec2 xx() const { return {x, x}; }
    vec2 yz() const { return {y, z}; }
    vec3& operator+=(vec3 b) { x += b.x; y += b.y; z += b.z; return *this; }
    vec3& operator-=(vec3 b) { x -= b.x; y -= b.y; z -= b.z; return *this; }
    vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
};
inline ivec3::ivec3(const vec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
inline uvec3::uvec3(const vec3& v) : x((uint)v.x), y((uint)v.y), z((uint)v.z) {}

struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    vec3 xyz() const { return {x, y, z}; }
    vec3 rgb() const { return {x, y, z}; }
};
struct uvec4 {
    uint x, y, z, w;
    uvec3 xyz() const { return {x, y, z}; }
};

// ---- arithmetic (component-wise, binary32, source order) ----
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator+(float s, vec3 a) { return {s + a.x, s + a.y, s + a.z}; }
inline vec3 operator-(float s, vec3 a) { return {s - a.x, s - a.y, s - a.z}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(float s, vec3 a) { return {s / a.x, s / a.y, s / a.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }

inline uvec3 operator-(uvec3 a, uvec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline uvec3 operator*(uvec3 a, uvec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline uvec3 operator&(uvec3 a, uvec3 b) { return {a.x & b.x, a.y & b.y, a.z & b.z}; }
inline uvec3 operator^(uvec3 a, uvec3 b) { return {a.x ^ b.x, a.y ^ b.y, a.z ^ b.z}; }
inline uint shr(uint v, uint s) { return s >= 32u ? 0u : v >> s; }
inline uint shl(uint v, uint s) { return s >= 32u ? 0u : v << s; }
inline uvec3 operator>>(uvec3 a, uint s) { return {shr(a.x, s), shr(a.y, s), shr(a.z, s)}; }
inline uvec3 operator<<(uvec3 a, uint s) { return {shl(a.x, s), shl(a.y, s), shl(a.z, s)}; }

// ---- built-in functions ----
inline float abs(float a) { return std::fabs(a); }
inline float floor(float a) { return std::floor(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float exp2(float a) { return std::exp2(a); } // only called with small negative integers: exact
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline float sign(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }
inline float mod(float x, float y) { return x - y * std::floor(x / y); }

inline vec3 abs(vec3 a) { return {abs(a.x), abs(a.y), abs(a.z)}; }
inline vec3 floor(vec3 a) { return {floor(a.x), floor(a.y), floor(a.z)}; }
inline vec3 sign(vec3 a) { return {sign(a.x), sign(a.y), sign(a.z)}; }
inline vec3 min(vec3 a, vec3 b) { return {min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)}; }
inline vec3 max(vec3 a, vec3 b) { return {max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)}; }
inline vec2 min(vec2 a, vec2 b) { return {min(a.x, b.x), min(a.y, b.y)}; }
inline vec2 max(vec2 a, vec2 b) { return {max(a.x, b.x), max(a.y, b.y)}; }
inline vec3 mod(vec3 a, float y) { return {mod(a.x, y), mod(a.y, y), mod(a.z, y)}; }

inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
inline vec3 normalize(vec3 a) { return a / sqrt(dot(a, a)); }

inline bvec3 lessThan(vec3 a, vec3 b) { return {a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bvec3 lessThanEqual(vec3 a, vec3 b) { return {a.x <= b.x, a.y <= b.y, a.z <= b.z}; }
inline bvec3 greaterThan(vec3 a, vec3 b) { return {a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bvec3 greaterThanEqual(vec3 a, vec3 b) { return {a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline bvec2 greaterThanEqual(uvec2 a, uvec2 b) { return {a.x >= b.x, a.y >= b.y}; }
inline bvec3 notEqual(uvec3 a, uvec3 b) { return {a.x != b.x, a.y != b.y, a.z != b.z}; }
inline bool any(bvec2 b) { return b.x || b.y; }

// mix(x, y, a) with boolean a selects y where a is true
inline vec3 mix(vec3 a, vec3 b, bvec3 c) { return {c.x ? b.x : a.x, c.y ? b.y : a.y, c.z ? b.z : a.z}; }
inline ivec3 mix(ivec3 a, ivec3 b, bvec3 c) { return {c.x ? b.x : a.x, c.y ? b.y : a.y, c.z ? b.z : a.z}; }
inline int mix(int a, int b, bool c) { return c ? b : a; }

inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uvec3 floatBitsToUint(vec3 v) { return {floatBitsToUint(v.x), floatBitsToUint(v.y), floatBitsToUint(v.z)}; }
inline vec3 uintBitsToFloat(uvec3 v) { return {uintBitsToFloat(v.x), uintBitsToFloat(v.y), uintBitsToFloat(v.z)}; }

inline vec4 unpackUnorm4x8(uint p) {
    return {(float)(p & 0xFFu) / 255.0f, (float)((p >> 8) & 0xFFu) / 255.0f,
            (float)((p >> 16) & 0xFFu) / 255.0f, (float)((p >> 24) & 0xFFu) / 255.0f};
}

// ---- opaque types ----
struct sampler3D {
    const uint8_t* texels = nullptr; // RGBA8, x fastest (tight BufferImageCopy, DdaRaytraceAlgorithm.cpp:36-47)
    uint64_t nx = 0, ny = 0, nz = 0;
};
inline vec4 texelFetch(const sampler3D& s, ivec3 p, int) {
    if (p.x < 0 || p.y < 0 || p.z < 0 || (uint64_t)p.x >= s.nx || (uint64_t)p.y >= s.ny || (uint64_t)p.z >= s.nz)
        return {0.f, 0.f, 0.f, 0.f};
    const uint8_t* t = s.texels + 4 * ((uint64_t)p.x + (uint64_t)p.y * s.nx + (uint64_t)p.z * s.nx * s.ny);
    return {(float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f, (float)t[3] / 255.0f};
}
inline ivec3 textureSize(const sampler3D& s, int) { return {(int)s.nx, (int)s.ny, (int)s.nz}; }

struct image2D {
    uint32_t* pixels = nullptr;
    uint32_t width = 0, height = 0;
};
inline uint32_t unorm8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v > 1.0f) v = 1.0f;
    return (uint32_t)std::nearbyint(v * 255.0f); // default rounding mode: half to even
}
inline void imageStore(image2D& img, ivec2 p, vec4 v) {
    if (p.x < 0 || p.y < 0 || (uint32_t)p.x >= img.width || (uint32_t)p.y >= img.height) return;
    img.pixels[(size_t)p.y * img.width + (size_t)p.x] =
        unorm8(v.x) | (unorm8(v.y) << 8) | (unorm8(v.z) << 16) | (unorm8(v.w) << 24);
}

// index guard for the ESVO stack read with an underflowed `scale` (UB in GLSL,
// esvo.comp:119-123; the loop exits right after, so the value read is never used)
inline uint xn_guard(uint i) { return i > 23u ? 23u : i; }

} // namespace glsl
