// xn_brick.h -- the bricked (Morton-in-brick) residency layout of RGBA8 grids in HBM.
//
// The reference uploads its grid into an "optimal tiling" 3-D image
// (src/render/DdaRaytraceAlgorithm.cpp:16-34) and leaves the element order to the driver.
// Here the order is explicit: voxels are grouped in 8x8x8 bricks (2 KiB); inside a brick the
// three low bits of x, y, z are bit-interleaved (x0 y0 z0 x1 y1 z1 x2 y2 z2), so one 32-byte
// sector is a 2x2x2 voxel cube and one 128-byte line a 4x4x2 block whatever the ray direction.
// Above the brick every axis owns one contiguous bit field of the index:
//
//   index = m3(x&7) | m3(y&7) << 1 | m3(z&7) << 2
//         | (x>>3) << hs[0] | (y>>3) << hs[1] | (z>>3) << hs[2]
//
// The two lower fields are padded to a power-of-two number of bricks; the top field is not, so
// the axis whose padding would cost most is put on top.  Every axis therefore has a bit MASK of
// the index bits it owns, and a DDA step of +-1 along an axis is
//
//   d = (d + K) & mask,   K = -mask for +1 (the carry runs through the foreign bits), -1 for -1,
//
// on that axis' own dilated coordinate d; the voxel index is d_x | d_y | d_z.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define XN_BRICK_HD __host__ __device__ __forceinline__
#else
#define XN_BRICK_HD inline
#endif

namespace xn {

struct BrickLayout {
    uint32_t hs[3];   // bit position of the brick-index field of x, y, z
    uint32_t fb[3];   // width of that field (the top axis' field runs to the top of the index)
    uint32_t top;     // axis that owns the highest field (not padded to a power of two)
    uint64_t mask[3]; // index bits owned by x, y, z
    uint64_t total;   // number of voxel slots including padding
};

XN_BRICK_HD uint32_t brick_m3(uint32_t a) { return (a & 1u) | ((a & 2u) << 2) | ((a & 4u) << 4); }
XN_BRICK_HD uint32_t brick_c3(uint64_t v) {
    return (uint32_t)((v & 1u) | ((v >> 2) & 2u) | ((v >> 4) & 4u));
}

// dilated coordinate of c on `axis`; consistent with the step arithmetic for the out-of-range
// values -1 and n (they wrap inside the axis' own bits)
XN_BRICK_HD uint64_t brick_axis(const BrickLayout& L, int axis, int64_t c) {
    const uint64_t lo = (uint64_t)brick_m3((uint32_t)c & 7u) << axis;
    const uint64_t hi = ((uint64_t)(c >> 3) << L.hs[axis]);
    return (lo | hi) & L.mask[axis];
}
XN_BRICK_HD uint64_t brick_index(const BrickLayout& L, uint32_t x, uint32_t y, uint32_t z) {
    return brick_axis(L, 0, x) | brick_axis(L, 1, y) | brick_axis(L, 2, z);
}
// coordinate on `axis` of index i
XN_BRICK_HD uint32_t brick_coord(const BrickLayout& L, int axis, uint64_t i) {
    const uint64_t field = (i & L.mask[axis]) >> L.hs[axis];
    return (uint32_t)(field << 3) | brick_c3(i >> axis);
}

inline uint32_t ceil_log2_u64(uint64_t n) {
    uint32_t b = 0;
    while ((1ull << b) < n) ++b;
    return b;
}

// top < 0: pick the axis whose power-of-two padding would waste most
inline BrickLayout make_brick_layout(uint64_t nx, uint64_t ny, uint64_t nz, int top = -1) {
    const uint64_t n[3] = {nx, ny, nz};
    uint64_t nb[3];
    uint32_t f[3];
    for (int a = 0; a < 3; ++a) {
        nb[a] = (n[a] + 7) / 8;
        f[a] = ceil_log2_u64(nb[a]);
    }
    if (top < 0) {
        double worst = -1.0;
        for (int a = 2; a >= 0; --a) { // ties go to z
            const double waste = (double)(1ull << f[a]) / (double)nb[a];
            if (waste > worst + 1e-12) {
                worst = waste;
                top = a;
            }
        }
    }
    BrickLayout L{};
    L.top = (uint32_t)top;
    uint32_t pos = 9;
    for (int a = 0; a < 3; ++a) {
        if (a == top) continue;
        L.hs[a] = pos;
        L.fb[a] = f[a];
        L.mask[a] = ((uint64_t)0x49u << a) | ((((uint64_t)1 << f[a]) - 1) << pos);
        pos += f[a];
    }
    L.hs[top] = pos;
    L.fb[top] = 64 - pos;
    L.mask[top] = ((uint64_t)0x49u << top) | (~(uint64_t)0 << pos);
    L.total = (nb[top] << pos);
    return L;
}

} // namespace xn
