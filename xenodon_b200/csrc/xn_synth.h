// xn_synth.h -- deterministic synthetic volumes, integer arithmetic only, so the host
// generator (xn_synth_grid_host) and the device generator (synth_kernel) produce
// bit-identical voxels.  Shapes follow the reference's data tools:
//   kind 0 "bunny-CT": grey solid with internal texture inside a cylinder mask, zero
//          elsewhere (tools/make-bunny-volume.py:29-48: threshold > 5, cylinder r = 250/512)
//   kind 1 "TNG gas" : filamentary ridged-noise density through a magma-like colour map,
//          ~90 % of voxels at the map's floor colour (tools/make-tng-volume.py:36-50)
// All voxels have alpha 255 (both tools write opaque RGBA).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define XN_HD __host__ __device__ inline
#else
#define XN_HD inline
#endif

namespace xn {

struct SynthSpec {
    uint32_t kind, nx, ny, nz, seed;
};

XN_HD uint32_t synth_rotl(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

XN_HD uint32_t synth_hash(uint32_t x, uint32_t y, uint32_t z, uint32_t seed) {
    uint32_t h = seed * 0x9E3779B1u + 0x7F4A7C15u;
    h ^= x * 0x85EBCA6Bu;
    h = synth_rotl(h, 15) * 0xC2B2AE35u;
    h ^= y * 0x27D4EB2Fu;
    h = synth_rotl(h, 13) * 0x165667B1u;
    h ^= z * 0x9E3779B1u;
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

XN_HD int32_t synth_lerp(int32_t a, int32_t b, uint32_t f16) {
    return a + (int32_t)(((int64_t)(b - a) * (int64_t)f16) >> 16);
}

// trilinear value noise on a lattice of spacing 2^shift voxels; result in [0, 65535]
XN_HD uint32_t synth_value_noise(uint32_t x, uint32_t y, uint32_t z, int shift, uint32_t seed) {
    if (shift <= 0) return synth_hash(x, y, z, seed) & 0xFFFFu;
    const uint32_t mask = (1u << shift) - 1u;
    const uint32_t i = x >> shift, j = y >> shift, k = z >> shift;
    const uint32_t fx = (x & mask) << (16 - shift), fy = (y & mask) << (16 - shift), fz = (z & mask) << (16 - shift);
    int32_t c[8];
    for (int d = 0; d < 8; ++d)
        c[d] = (int32_t)(synth_hash(i + ((d >> 2) & 1), j + ((d >> 1) & 1), k + (d & 1), seed) & 0xFFFFu);
    const int32_t x00 = synth_lerp(c[0], c[4], fx), x01 = synth_lerp(c[1], c[5], fx);
    const int32_t x10 = synth_lerp(c[2], c[6], fx), x11 = synth_lerp(c[3], c[7], fx);
    const int32_t y0 = synth_lerp(x00, x10, fy), y1 = synth_lerp(x01, x11, fy);
    return (uint32_t)synth_lerp(y0, y1, fz);
}

XN_HD int synth_ilog2(uint32_t v) {
    int l = 0;
    while (v >>= 1) ++l;
    return l;
}

// inside test for an axis-aligned ellipsoid, coordinates in 1/65536 of the box
XN_HD bool synth_in_ellipsoid(int32_t cx, int32_t cy, int32_t cz, int32_t ex, int32_t ey, int32_t ez, int32_t rx,
                              int32_t ry, int32_t rz) {
    const int64_t qx = ((int64_t)(cx - ex) * 1024) / rx;
    const int64_t qy = ((int64_t)(cy - ey) * 1024) / ry;
    const int64_t qz = ((int64_t)(cz - ez) * 1024) / rz;
    return qx * qx + qy * qy + qz * qz < (int64_t)1024 * 1024;
}

XN_HD uint32_t synth_bunny(const SynthSpec& s, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t opaque_black = 0xFF000000u;
    const int32_t cx = (int32_t)(((uint64_t)x * 65536u + 32768u) / s.nx);
    const int32_t cy = (int32_t)(((uint64_t)y * 65536u + 32768u) / s.ny);
    const int32_t cz = (int32_t)(((uint64_t)z * 65536u + 32768u) / s.nz);
    // cylinder mask around the y axis, radius 250/512 of the box
    const int64_t dx = cx - 32768, dz = cz - 32768;
    if (dx * dx + dz * dz > (int64_t)32000 * 32000) return opaque_black;
    const bool solid = synth_in_ellipsoid(cx, cy, cz, 32768, 26000, 31000, 22000, 19000, 25000) ||
                       synth_in_ellipsoid(cx, cy, cz, 32768, 46000, 41000, 12500, 11500, 13000) ||
                       synth_in_ellipsoid(cx, cy, cz, 27500, 57500, 40000, 3200, 7000, 4200) ||
                       synth_in_ellipsoid(cx, cy, cz, 38000, 57500, 40000, 3200, 7000, 4200);
    if (!solid) return opaque_black;
    const uint32_t n = synth_value_noise(x, y, z, 3, s.seed);
    int32_t v = (int32_t)((200u * (154u + ((102u * n) >> 16))) >> 8);
    v += (int32_t)(synth_hash(x, y, z, s.seed ^ 0xA511E9B3u) % 13u) - 6;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    if (v <= 5) v = 0;
    return opaque_black | (uint32_t)v | ((uint32_t)v << 8) | ((uint32_t)v << 16);
}

// magma-like colour map: piecewise-linear through five control points
XN_HD uint32_t synth_magma(uint32_t t) {
    const int32_t cp[5][3] = {{0, 0, 3}, {80, 18, 123}, {182, 54, 121}, {251, 136, 97}, {252, 253, 191}};
    const uint32_t seg = t >> 6, f = (t & 63u) << 10; // 4 segments of 64 steps, 16-bit fraction
    const uint32_t hi = seg + 1u > 4u ? 4u : seg + 1u;
    const uint32_t r = (uint32_t)synth_lerp(cp[seg][0], cp[hi][0], f);
    const uint32_t g = (uint32_t)synth_lerp(cp[seg][1], cp[hi][1], f);
    const uint32_t b = (uint32_t)synth_lerp(cp[seg][2], cp[hi][2], f);
    return 0xFF000000u | r | (g << 8) | (b << 16);
}

XN_HD uint32_t synth_ridged(const SynthSpec& s, uint32_t x, uint32_t y, uint32_t z, int s0, uint32_t seed) {
    const uint32_t o0 = synth_value_noise(x, y, z, s0, seed);
    const uint32_t o1 = synth_value_noise(x, y, z, s0 - 1, seed + 1u);
    const uint32_t o2 = synth_value_noise(x, y, z, s0 - 2, seed + 2u);
    const int32_t n = (int32_t)((4u * o0 + 2u * o1 + o2) / 7u); // [0, 65535]
    int32_t d = n - 32768;
    d = d < 0 ? -d : d;
    const int32_t r = 65535 - 4 * d; // ridge along the noise's mid level set
    (void)s;
    return (uint32_t)(r < 0 ? 0 : r);
}

XN_HD uint32_t synth_tng(const SynthSpec& s, uint32_t x, uint32_t y, uint32_t z) {
    uint32_t maxdim = s.nx > s.ny ? s.nx : s.ny;
    maxdim = maxdim > s.nz ? maxdim : s.nz;
    const int s0 = synth_ilog2(maxdim) - 3; // coarsest lattice = 1/8 of the box
    // intersection of two ridged fields -> filaments
    const uint32_t r1 = synth_ridged(s, x, y, z, s0, s.seed);
    const uint32_t r2 = synth_ridged(s, x, y, z, s0, s.seed + 101u);
    const uint32_t r = r1 < r2 ? r1 : r2;
    // tools/make-tng-volume.py normalises log-density between a low percentile (90th for the
    // 1024^3 volume, 96th for 2048^3) and the 99th percentile and clips; the constants below are
    // those percentiles of this ridge field (measured once; the field's distribution does not
    // depend on the volume size), so ~90 % / ~96 % of voxels sit at the colour map's floor.
    const uint32_t lo = maxdim >= 2048u ? 58410u : 54150u;
    const uint32_t hi = 61960u;
    if (r <= lo) return synth_magma(0);
    const uint32_t t = ((r - lo) * 255u) / (hi - lo);
    return synth_magma(t > 255u ? 255u : t);
}

XN_HD uint32_t synth_voxel(const SynthSpec& s, uint32_t x, uint32_t y, uint32_t z) {
    return s.kind == 0 ? synth_bunny(s, x, y, z) : synth_tng(s, x, y, z);
}

} // namespace xn
