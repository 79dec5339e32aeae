// xn_kernels.cu -- the five volume-traversal kernels for sm_100a.  Compiled with -fmad=false
// (see xn_device.cuh): FMAs appear only where written explicitly.
//
// Each kernel restates, from scratch, what one reference compute shader computes:
//   dda_kernel        resources/dda.comp        Amanatides-Woo grid march (multi-axis tie steps)
//   svo_naive_kernel  resources/svo_naive.comp  root-restart point location per leaf
//   svo_df_kernel     resources/svo_df.comp     exhaustive depth-first visit, explicit stack
//   esvo_kernel       resources/esvo.comp       Laine-Karras ESVO with emission accumulation
//   svo_rope_kernel   resources/svo_rope.comp   rope-tree leaf-to-leaf walk
// One thread per pixel; a warp owns an 8x4 pixel tile.  Traversal stacks live in shared
// memory ([level][thread], conflict-free).
//
// Two arithmetic modes (template STRICT):
//   STRICT : every operation, including colour accumulation, in the shader's order -> images
//            bit-identical to the CPU oracle.
//   fast   : the GEOMETRY (which voxels / nodes are visited, every segment length) is still
//            computed in the shader's exact order, so per-ray step counts are identical; only
//            the colour sum is accumulated as fma(byte, length, sum) and scaled by 1/255 once
//            per ray.  Differs from STRICT by at most 1/255 on a small fraction of pixels.
#include <cstdlib>

#include "xn_brick.h"
#include "xn_device.cuh"
#include "xn_kernels.h"

// tuning switches (A/B-tested on the B200, see profiles/): how many of the three colour
// channels of the fast mode convert byte -> float with I2F.U8 rather than PRMT + FADD
#ifndef XN_FAST_I2F
#define XN_FAST_I2F 3
#endif
// octree kernels: minimum resident blocks per SM requested from the register allocator
#ifndef XN_SVO_MIN_BLOCKS
#define XN_SVO_MIN_BLOCKS 1
#endif
// DDA: skip the colour arithmetic of a trip whose texels are black across the whole warp
#ifndef XN_DDA_SKIP_EMPTY
#define XN_DDA_SKIP_EMPTY 1
#endif
// DDA: march long in-grid stretches as unchecked segments (no per-step bounds test / position)
#ifndef XN_DDA_SEGMENTS
#define XN_DDA_SEGMENTS 1
#endif
// octree descent (svo_naive / svo_rope): child selection with comparison flags and FMAs
#ifndef XN_DESCEND_FLAGS
#define XN_DESCEND_FLAGS 1
#endif
// minimum resident blocks requested for the ESVO (39 registers at 6: +1.3 %) and the texture DDA
#ifndef XN_ESVO_MIN_BLOCKS
#define XN_ESVO_MIN_BLOCKS 6
#endif
#ifndef XN_TEX_MIN_BLOCKS
#define XN_TEX_MIN_BLOCKS 6
#endif
// svo_df: hit mask of a node accumulated with comparison flags and FMAs
#ifndef XN_DF_FLAG_MASK
#define XN_DF_FLAG_MASK 1
#endif
// node slab test of svo_naive / svo_rope with FMNMX
#ifndef XN_SLAB_FMNMX
#define XN_SLAB_FMNMX 1
#endif
// svo_naive: first TOP_LEVELS levels of find() from the entry table
#ifndef XN_NAIVE_TOP_TABLE
#define XN_NAIVE_TOP_TABLE 1
#endif
// ESVO PUSH: 1 = always write the stack entry, 0 = only when the child exits before its parent (`h`)
#ifndef XN_ESVO_ALWAYS_STORE
#define XN_ESVO_ALWAYS_STORE 1
#endif
// ESVO POP: 1 = differing bits from the old/new positions and mask-based truncation
#ifndef XN_ESVO_POP
#define XN_ESVO_POP 1
#endif

namespace xn {

template <bool STATS>
struct RayStats {
    uint32_t steps = 0;
    unsigned long long bytes = 0;
    __device__ __forceinline__ void step() {
        if (STATS) ++steps;
    }
    __device__ __forceinline__ void read(uint32_t n) {
        if (STATS) bytes += n;
    }
};

// One 256-bit read-only load of a 32-byte, 32-byte-aligned record (sm_100: LDG.E.256): one request
// and one sector per lane where two 128-bit loads are two requests for the same sector.
#ifndef XN_LDG256
#define XN_LDG256 1
#endif
__device__ __forceinline__ void ldg256(const void* ptr, uint4& lo, uint4& hi) {
    asm("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
        : "l"(ptr));
}

// (float)byte k of v, exactly, on the ALU/FMA pipes: 0x4B0000bb is 2^23 + bb
template <int K>
__device__ __forceinline__ float byte_f(uint32_t v) {
    return __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7650u | K)) - 8388608.0f;
}

// exact (float)b / 255.0f without an IEEE division: multiply by the rounded reciprocal plus one
// FMA Newton correction (all 256 inputs verified, tests/test_host_formats.py)
__device__ __forceinline__ float unorm8_exact(float x) {
    const float r = 1.0f / 255.0f;
    const float q = x * r;
    const float rem = __fmaf_rn(-q, 255.0f, x);
    return __fmaf_rn(rem, r, q);
}

// emission accumulator: sum of colour * segment length
template <bool STRICT>
struct Accum {
    float r = 0.f, g = 0.f, b = 0.f;
    // rgb = packed r | g<<8 | b<<16 (texel or node colour), len = segment length
    __device__ __forceinline__ void add(uint32_t rgb, float len) {
        if (STRICT) {
            r += unorm8_exact(byte_f<0>(rgb)) * len;
            g += unorm8_exact(byte_f<1>(rgb)) * len;
            b += unorm8_exact(byte_f<2>(rgb)) * len;
        } else {
            // XN_FAST_I2F of the three channels convert with I2F.U8 (one instruction on the
            // quarter-rate conversion pipe), the rest with PRMT + FADD (ALU + FMA pipes)
            r = __fmaf_rn(XN_FAST_I2F >= 1 ? (float)(rgb & 0xFFu) : byte_f<0>(rgb), len, r);
            g = __fmaf_rn(XN_FAST_I2F >= 2 ? (float)((rgb >> 8) & 0xFFu) : byte_f<1>(rgb), len, g);
            b = __fmaf_rn(XN_FAST_I2F >= 3 ? (float)((rgb >> 16) & 0xFFu) : byte_f<2>(rgb), len, b);
        }
    }
    __device__ __forceinline__ f3 finish(float ec) const {
        const float s = STRICT ? ec : ec / 255.0f;
        return F3(r * s, g * s, b * s);
    }
};

template <bool STATS>
__device__ __forceinline__ void store_result(const FrameParams& p, uint32_t ix, uint32_t iy, f3 color,
                                             const RayStats<STATS>& st) {
    p.target[(uint64_t)iy * p.target_stride + ix] = pack_pixel(color);
    if (STATS) {
        const uint64_t i = (uint64_t)iy * p.out_w + ix;
        if (p.steps_out) p.steps_out[i] = st.steps;
        if (p.bytes_out) p.bytes_out[i] = st.bytes;
    }
}

// Position of a ray inside the grid as an address.  The voxel coordinates themselves are only
// tracked in the checked single steps; unchecked segments move the cursor alone and recover the
// coordinates from it afterwards.
//   GridCursor<false>: one signed 32-bit voxel index (grids below 2^31 voxels).
//   GridCursor<true> : a 64-bit pointer to the current z slice plus a 32-bit index inside the
//                      slice, so that only z steps pay 64-bit arithmetic (2048^3 = 2^33 voxels).
template <bool BIG>
struct GridCursor;

template <>
struct GridCursor<false> {
    const uint32_t* __restrict__ grid;
    int32_t idx, dix, diy, diz, stride_y, stride_z;
    __device__ __forceinline__ GridCursor(const FrameParams& p, int px, int py, int pz, int sx, int sy, int sz)
        : grid(p.grid), stride_y((int32_t)p.nx), stride_z((int32_t)(p.nx * p.ny)) {
        dix = sx;
        diy = sy * stride_y;
        diz = sz * stride_z;
        idx = px + py * stride_y + pz * stride_z;
    }
    __device__ __forceinline__ void step_x() { idx += dix; }
    __device__ __forceinline__ void step_y() { idx += diy; }
    __device__ __forceinline__ void step_z() { idx += diz; }
    __device__ __forceinline__ uint32_t load() const { return __ldg(grid + idx); }
    // voxel coordinates of the (in-range) cursor
    __device__ __forceinline__ void recover(int& px, int& py, int& pz) const {
        const uint32_t qz = (uint32_t)idx / (uint32_t)stride_z;
        const uint32_t rem = (uint32_t)idx - qz * (uint32_t)stride_z;
        const uint32_t qy = rem / (uint32_t)stride_y;
        pz = (int)qz;
        py = (int)qy;
        px = (int)(rem - qy * (uint32_t)stride_y);
    }
};

template <>
struct GridCursor<true> {
    const uint32_t* __restrict__ grid;
    const uint32_t* slice; // grid + pz * stride_z (may point outside the grid while out of range)
    int64_t diz, stride_z;
    int32_t ixy, dix, diy, stride_y;
    __device__ __forceinline__ GridCursor(const FrameParams& p, int px, int py, int pz, int sx, int sy, int sz)
        : grid(p.grid), stride_z((int64_t)p.nx * (int64_t)p.ny), stride_y((int32_t)p.nx) {
        dix = sx;
        diy = sy * stride_y;
        diz = (int64_t)sz * stride_z;
        ixy = px + py * stride_y;
        slice = grid + (int64_t)pz * stride_z;
    }
    __device__ __forceinline__ void step_x() { ixy += dix; }
    __device__ __forceinline__ void step_y() { ixy += diy; }
    __device__ __forceinline__ void step_z() { slice += diz; }
    __device__ __forceinline__ uint32_t load() const { return __ldg(slice + ixy); }
    __device__ __forceinline__ void recover(int& px, int& py, int& pz) const {
        // slice offset < 2^48: the binary64 quotient is off by at most one, fixed below
        const int64_t off = (int64_t)(slice - grid);
        int64_t qz = (int64_t)((double)off / (double)stride_z);
        const int64_t rem = off - qz * stride_z;
        if (rem < 0) --qz;
        else if (rem >= stride_z) ++qz;
        const uint32_t qy = (uint32_t)ixy / (uint32_t)stride_y;
        pz = (int)qz;
        py = (int)qy;
        px = (int)((uint32_t)ixy - qy * (uint32_t)stride_y);
    }
};

// Cursor over the bricked layout (xn_brick.h): every axis keeps its own dilated coordinate, a
// step is one add and one mask on that coordinate, the voxel index is the OR of the three.
//   BrickCursor<false>: 32-bit index (up to 2^32 voxel slots, any axis on top).
//   BrickCursor<true> : z on top with a 64-bit dilated coordinate; x and y stay 32-bit.
template <bool BIG>
struct BrickCursor;

// a | b | c as ONE LOP3 (the compiler otherwise emits two two-input ORs per texel address)
__device__ __forceinline__ uint32_t or3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xFE;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <>
struct BrickCursor<false> {
    const uint32_t* __restrict__ grid;
    uint32_t dx, dy, dz, kx, ky, kz, mx, my, mz, hx, hy, hz;
    static __device__ __forceinline__ uint32_t enc(int c, int axis, uint32_t hs, uint32_t mask) {
        return ((brick_m3((uint32_t)c & 7u) << axis) | ((uint32_t)(c >> 3) << hs)) & mask;
    }
    static __device__ __forceinline__ int dec(uint32_t d, int axis, uint32_t hs) {
        return (int)(((d >> hs) << 3) | brick_c3(d >> axis));
    }
    __device__ __forceinline__ BrickCursor(const FrameParams& p, int px, int py, int pz, int sx, int sy, int sz)
        : grid(p.grid), mx(p.bk_mask[0]), my(p.bk_mask[1]), mz(p.bk_mask[2]), hx(p.bk_hs[0]), hy(p.bk_hs[1]),
          hz(p.bk_hs[2]) {
        dx = enc(px, 0, hx, mx);
        dy = enc(py, 1, hy, my);
        dz = enc(pz, 2, hz, mz);
        kx = sx > 0 ? 0u - mx : 0xFFFFFFFFu;
        ky = sy > 0 ? 0u - my : 0xFFFFFFFFu;
        kz = sz > 0 ? 0u - mz : 0xFFFFFFFFu;
    }
    __device__ __forceinline__ void step_x() { dx = (dx + kx) & mx; }
    __device__ __forceinline__ void step_y() { dy = (dy + ky) & my; }
    __device__ __forceinline__ void step_z() { dz = (dz + kz) & mz; }
    __device__ __forceinline__ uint32_t load() const { return __ldg(grid + or3(dx, dy, dz)); }
    __device__ __forceinline__ void recover(int& px, int& py, int& pz) const {
        px = dec(dx, 0, hx);
        py = dec(dy, 1, hy);
        pz = dec(dz, 2, hz);
    }
};

template <>
struct BrickCursor<true> {
    const uint32_t* __restrict__ grid;
    uint64_t dz, kz, mz;
    uint32_t dx, dy, kx, ky, mx, my, hx, hy, hz;
    __device__ __forceinline__ BrickCursor(const FrameParams& p, int px, int py, int pz, int sx, int sy, int sz)
        : grid(p.grid), mz(p.bk_mask_z64), mx(p.bk_mask[0]), my(p.bk_mask[1]), hx(p.bk_hs[0]), hy(p.bk_hs[1]),
          hz(p.bk_hs[2]) {
        dx = BrickCursor<false>::enc(px, 0, hx, mx);
        dy = BrickCursor<false>::enc(py, 1, hy, my);
        dz = (((uint64_t)brick_m3((uint32_t)pz & 7u) << 2) | ((uint64_t)(int64_t)(pz >> 3) << hz)) & mz;
        kx = sx > 0 ? 0u - mx : 0xFFFFFFFFu;
        ky = sy > 0 ? 0u - my : 0xFFFFFFFFu;
        kz = sz > 0 ? 0ull - mz : ~0ull;
    }
    __device__ __forceinline__ void step_x() { dx = (dx + kx) & mx; }
    __device__ __forceinline__ void step_y() { dy = (dy + ky) & my; }
    __device__ __forceinline__ void step_z() { dz = (dz + kz) & mz; }
    __device__ __forceinline__ uint32_t load() const {
        return __ldg(grid + (((dz >> 32) << 32) | (uint64_t)or3(dx, dy, (uint32_t)dz)));
    }
    __device__ __forceinline__ void recover(int& px, int& py, int& pz) const {
        px = BrickCursor<false>::dec(dx, 0, hx);
        py = BrickCursor<false>::dec(dy, 1, hy);
        pz = (int)(((dz >> hz) << 3) | brick_c3(dz >> 2));
    }
};

// One DDA step (dda.comp:41-50): t0 = min(side distances), dt = t0 - t, every axis whose side
// distance equals t0 steps.  Generic form in C++; for the 32-bit linear cursor the step is written
// in PTX with predicated adds, because the compiler otherwise turns the three conditional index
// updates into SEL + three-input adds -- five ALU-pipe instructions on the pipe that bounds this
// kernel (ncu: ALU 69 %) instead of three predicated adds.
#ifndef XN_DDA_PTX_STEP
#define XN_DDA_PTX_STEP 1
#endif
template <class CURSOR>
__device__ __forceinline__ void dda_step(CURSOR& cur, float& sdx, float& sdy, float& sdz, float tdx, float tdy,
                                         float tdz, float& t, float& dt) {
    const float t0 = fminf(sdx, fminf(sdy, sdz));
    const bool mx = sdx == t0, my = sdy == t0, mz = sdz == t0;
    dt = t0 - t;
    t = t0;
    if (mx) { sdx += tdx; cur.step_x(); }
    if (my) { sdy += tdy; cur.step_y(); }
    if (mz) { sdz += tdz; cur.step_z(); }
}
#if XN_DDA_PTX_STEP
template <>
__device__ __forceinline__ void dda_step<GridCursor<false>>(GridCursor<false>& cur, float& sdx, float& sdy, float& sdz,
                                                            float tdx, float tdy, float tdz, float& t, float& dt) {
    asm("{\n\t"
        ".reg .pred px, py, pz;\n\t"
        ".reg .f32 t0;\n\t"
        "min.f32 t0, %1, %2;\n\t"
        "min.f32 t0, %0, t0;\n\t"
        "setp.eq.f32 px, %0, t0;\n\t"
        "setp.eq.f32 py, %1, t0;\n\t"
        "setp.eq.f32 pz, %2, t0;\n\t"
        "sub.f32 %5, t0, %4;\n\t"
        "mov.f32 %4, t0;\n\t"
        "@px add.f32 %0, %0, %6;\n\t"
        "@py add.f32 %1, %1, %7;\n\t"
        "@pz add.f32 %2, %2, %8;\n\t"
        "@px add.s32 %3, %3, %9;\n\t"
        "@py add.s32 %3, %3, %10;\n\t"
        "@pz add.s32 %3, %3, %11;\n\t"
        "}"
        : "+f"(sdx), "+f"(sdy), "+f"(sdz), "+r"(cur.idx), "+f"(t), "=f"(dt)
        : "f"(tdx), "f"(tdy), "f"(tdz), "r"(cur.dix), "r"(cur.diy), "r"(cur.diz));
}
// the same for the bricked cursors: t = d + K is unconditional, the mask is applied under the
// predicate (d = t & mask), so an axis step is one IADD and one predicated LOP3
template <>
__device__ __forceinline__ void dda_step<BrickCursor<false>>(BrickCursor<false>& cur, float& sdx, float& sdy,
                                                             float& sdz, float tdx, float tdy, float tdz, float& t,
                                                             float& dt) {
    asm("{\n\t"
        ".reg .pred px, py, pz;\n\t"
        ".reg .f32 t0;\n\t"
        ".reg .b32 ux, uy, uz;\n\t"
        "min.f32 t0, %1, %2;\n\t"
        "min.f32 t0, %0, t0;\n\t"
        "setp.eq.f32 px, %0, t0;\n\t"
        "setp.eq.f32 py, %1, t0;\n\t"
        "setp.eq.f32 pz, %2, t0;\n\t"
        "sub.f32 %7, t0, %6;\n\t"
        "mov.f32 %6, t0;\n\t"
        "add.u32 ux, %3, %11;\n\t"
        "add.u32 uy, %4, %12;\n\t"
        "add.u32 uz, %5, %13;\n\t"
        "@px add.f32 %0, %0, %8;\n\t"
        "@py add.f32 %1, %1, %9;\n\t"
        "@pz add.f32 %2, %2, %10;\n\t"
        "@px and.b32 %3, ux, %14;\n\t"
        "@py and.b32 %4, uy, %15;\n\t"
        "@pz and.b32 %5, uz, %16;\n\t"
        "}"
        : "+f"(sdx), "+f"(sdy), "+f"(sdz), "+r"(cur.dx), "+r"(cur.dy), "+r"(cur.dz), "+f"(t), "=f"(dt)
        : "f"(tdx), "f"(tdy), "f"(tdz), "r"(cur.kx), "r"(cur.ky), "r"(cur.kz), "r"(cur.mx), "r"(cur.my), "r"(cur.mz));
}
template <>
__device__ __forceinline__ void dda_step<BrickCursor<true>>(BrickCursor<true>& cur, float& sdx, float& sdy, float& sdz,
                                                            float tdx, float tdy, float tdz, float& t, float& dt) {
    asm("{\n\t"
        ".reg .pred px, py, pz;\n\t"
        ".reg .f32 t0;\n\t"
        ".reg .b32 ux, uy;\n\t"
        ".reg .b64 uz;\n\t"
        "min.f32 t0, %1, %2;\n\t"
        "min.f32 t0, %0, t0;\n\t"
        "setp.eq.f32 px, %0, t0;\n\t"
        "setp.eq.f32 py, %1, t0;\n\t"
        "setp.eq.f32 pz, %2, t0;\n\t"
        "sub.f32 %7, t0, %6;\n\t"
        "mov.f32 %6, t0;\n\t"
        "add.u32 ux, %3, %11;\n\t"
        "add.u32 uy, %4, %12;\n\t"
        "add.u64 uz, %5, %13;\n\t"
        "@px add.f32 %0, %0, %8;\n\t"
        "@py add.f32 %1, %1, %9;\n\t"
        "@pz add.f32 %2, %2, %10;\n\t"
        "@px and.b32 %3, ux, %14;\n\t"
        "@py and.b32 %4, uy, %15;\n\t"
        "@pz and.b64 %5, uz, %16;\n\t"
        "}"
        : "+f"(sdx), "+f"(sdy), "+f"(sdz), "+r"(cur.dx), "+r"(cur.dy), "+l"(cur.dz), "+f"(t), "=f"(dt)
        : "f"(tdx), "f"(tdy), "f"(tdz), "r"(cur.kx), "r"(cur.ky), "l"(cur.kz), "r"(cur.mx), "r"(cur.my), "l"(cur.mz));
}
#endif

// ---------------------------------------------------------------------------------
// DDA (resources/dda.comp:13-73)
// CURSOR selects the resident layout and index width: GridCursor<false/true> = x-major linear
// (32-bit index / grids of 2^31 voxels and more), BrickCursor<false/true> = bricked (xn_brick.h).
// Texels are requested ahead of their accumulation (their addresses never depend on loaded
// data), so every warp overlaps its own load latency with arithmetic.
// ---------------------------------------------------------------------------------
template <bool STATS, bool STRICT, class CURSOR>
__global__ void __launch_bounds__(BLOCK_THREADS) dda_kernel(const __grid_constant__ FrameParams p) {
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    // textureSize(model) = grid dimensions; side = largest
    const float side = fmaxf((float)p.nx, fmaxf((float)p.ny, (float)p.nz));
    f3 ro = F3(p.pos[0] * side, p.pos[1] * side, p.pos[2] * side);
    const float ec = voxel_emission_coeff(p, rd) / side;

    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);
    const f3 bmin = F3(-bias.x, -bias.y, -bias.z);
    const f3 bmax = F3((float)p.model_dim[0] * rrd.x - bias.x, (float)p.model_dim[1] * rrd.y - bias.y,
                       (float)p.model_dim[2] * rrd.z - bias.z);
    float t_min = max_elem(F3(gmin(bmin.x, bmax.x), gmin(bmin.y, bmax.y), gmin(bmin.z, bmax.z)));
    const float t_max = min_elem(F3(gmax(bmin.x, bmax.x), gmax(bmin.y, bmax.y), gmax(bmin.z, bmax.z)));

    Accum<STRICT> acc;
    if (!(t_min > t_max)) {
        t_min = gmax(t_min, 0.0f);
        ro = F3(ro.x + rd.x * t_min, ro.y + rd.y * t_min, ro.z + rd.z * t_min);
        int px = (int)ro.x, py = (int)ro.y, pz = (int)ro.z; // ivec3(ro): truncation

        const float tdx = fabsf(rrd.x), tdy = fabsf(rrd.y), tdz = fabsf(rrd.z);
        const f3 sg = F3(gsign(rd.x), gsign(rd.y), gsign(rd.z));
        const int sx = (int)sg.x, sy = (int)sg.y, sz = (int)sg.z;
        float sdx = (sg.x * ((floorf(ro.x) - ro.x) + 0.5f) + 0.5f) * tdx;
        float sdy = (sg.y * ((floorf(ro.y) - ro.y) + 0.5f) + 0.5f) * tdy;
        float sdz = (sg.z * ((floorf(ro.z) - ro.z) + 0.5f) + 0.5f) * tdz;

        CURSOR cur(p, px, py, pz, sx, sy, sz);
        const bool skip_empty = XN_DDA_SKIP_EMPTY && p.skip_empty != 0u; // volume has black background

        // texelFetch; outside the grid -> 0 (border)
        bool inr = (uint32_t)px < p.nx && (uint32_t)py < p.ny && (uint32_t)pz < p.nz;
        uint32_t v = 0;
        if (inr) v = cur.load();

        float t = 0.0f;
        const float t_end = t_max - t_min;
        while (t < t_end) {
#if XN_DDA_SEGMENTS
            // Unchecked segment.  From an in-range voxel with r_i steps left to the grid face on
            // axis i, no axis can leave the grid within min(r_i) iterations (an iteration takes at
            // most one step per axis).  Axis i takes its j-th step in the iteration that ends at
            // side distance sd_i + (j-1)*td_i, so bounding t by min_i(sd_i + (r_i-2)*td_i) (one
            // step of slack against rounding of the repeated additions) keeps every fetch of the
            // segment inside the grid without tracking the position or testing bounds per step.
            // The arithmetic on t / side distances is exactly the loop body of dda.comp:41-50.
            if (inr) {
                const int rx = sx > 0 ? (int)p.nx - 1 - px : px;
                const int ry = sy > 0 ? (int)p.ny - 1 - py : py;
                const int rz = sz > 0 ? (int)p.nz - 1 - pz : pz;
                if (min(rx, min(ry, rz)) >= 4) {
                    // slack: two steps, plus one per 2048 steps of the segment -- the rounding of the
                    // repeated sd += td can drift by ~2^-25 N^2 td over N steps inside a binade, which
                    // on axes of 8192 voxels and more would exceed a fixed slack
                    const float t_lim = fminf(
                        t_end, fminf(sdx + (float)(rx - 2 - (rx >> 11)) * tdx,
                                     fminf(sdy + (float)(ry - 2 - (ry >> 11)) * tdy,
                                           sdz + (float)(rz - 2 - (rz >> 11)) * tdz)));
                    if (t < t_lim) {
                        // One step advances t by at most td_min (the axis with the smallest
                        // spacing crosses a face within td_min), so while t < t_lim - 4 td_min four
                        // more steps are certain to be taken: they run as one TRIP without any
                        // per-step loop test.  A trip advances the geometry of its four steps first
                        // (it never depends on loaded data), requests their four texels together,
                        // and only then accumulates the four texels requested by the PREVIOUS trip,
                        // so a load has a whole trip to arrive (two register sets, a / b).
                        const float t_lim4 = t_lim - 4.5f * fminf(tdx, fminf(tdy, tdz));
#define XN_DDA_STEP(DT)                                        \
    {                                                          \
        dda_step(cur, sdx, sdy, sdz, tdx, tdy, tdz, t, DT);    \
        st.step();                                             \
        st.read(4);                                            \
    }
#define XN_DDA_TRIP(N, P)                                                        \
    {                                                                            \
        float d0;                                                                \
        XN_DDA_STEP(d0) N##0 = cur.load();                                        \
        XN_DDA_STEP(N##d1) N##1 = cur.load();                                     \
        XN_DDA_STEP(N##d2) N##2 = cur.load();                                     \
        XN_DDA_STEP(N##d3) N##3 = cur.load();                                     \
        /* empty space: when the four pending texels are black in every lane of the warp  */ \
        /* the twelve conversions and twelve multiply-adds are skipped (adding 0 is exact) */ \
        if (!skip_empty ||                                                       \
            __any_sync(__activemask(), ((P##0 | P##1 | P##2 | P##3) & 0x00FFFFFFu) != 0u)) {   \
            acc.add(P##0, P##d1);                                                \
            acc.add(P##1, P##d2);                                                \
            acc.add(P##2, P##d3);                                                \
            acc.add(P##3, d0);                                                   \
        }                                                                        \
    }
                        if (t < t_lim4) {
                            // set a starts as "nothing pending" except the current voxel's texel
                            uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = v, b0, b1, b2, b3;
                            float ad1 = 0.f, ad2 = 0.f, ad3 = 0.f, bd1, bd2, bd3;
                            for (;;) {
                                XN_DDA_TRIP(b, a)
                                if (!(t < t_lim4)) {
                                    acc.add(b0, bd1);
                                    acc.add(b1, bd2);
                                    acc.add(b2, bd3);
                                    v = b3;
                                    break;
                                }
                                XN_DDA_TRIP(a, b)
                                if (!(t < t_lim4)) {
                                    acc.add(a0, ad1);
                                    acc.add(a1, ad2);
                                    acc.add(a2, ad3);
                                    v = a3;
                                    break;
                                }
                            }
                        }
#undef XN_DDA_TRIP
                        // the last few steps of the segment, one at a time
                        while (t < t_lim) {
                            float dt;
                            XN_DDA_STEP(dt)
                            const uint32_t vn = cur.load();
                            acc.add(v, dt);
                            v = vn;
                        }
#undef XN_DDA_STEP
                        // recover the voxel coordinates from the linear index (still in range)
                        cur.recover(px, py, pz);
                        continue;
                    }
                }
            }
#endif
            // checked single step (grid faces, entry ties, the possible extra last iteration)
            // side distances stay finite and non-negative, so fminf == GLSL min, and
            // `sd.x <= min(sd.y, sd.z)` (dda.comp:42) <=> `sd.x == min(sd.x, sd.y, sd.z)`
            const float t0 = fminf(sdx, fminf(sdy, sdz));
            const bool mx = sdx == t0;
            const bool my = sdy == t0;
            const bool mz = sdz == t0;
            const float dt = t0 - t;
            t = t0;
            if (mx) { sdx += tdx; px += sx; cur.step_x(); }
            if (my) { sdy += tdy; py += sy; cur.step_y(); }
            if (mz) { sdz += tdz; pz += sz; cur.step_z(); }

            inr = (uint32_t)px < p.nx && (uint32_t)py < p.ny && (uint32_t)pz < p.nz;
            uint32_t vn = 0; // texel of the next step (used only if the loop continues)
            if (inr) vn = cur.load();

            acc.add(v, dt);
            v = vn;
            st.step();
            st.read(4);
        }
    }
    store_result(p, ix, iy, acc.finish(ec), st);
}

// ---------------------------------------------------------------------------------
// DDA over the texture residency (resources/dda.comp:13-73 again, same arithmetic).
// The grid lives in a 3-D CUDA array; the texture unit does what the cursor code does for
// the other layouts: address arithmetic (block-linear tiling, so a warp's texels share
// sectors whatever the ray direction), the bounds test (border colour 0 = the reference's
// clamp-to-border sampler, src/render/DdaRaytraceAlgorithm.cpp:26-29) and, in the fast
// mode, the byte -> float conversion.  The voxel position is kept as three floats at texel
// centres, stepped with predicated FADDs like the side distances.
// ---------------------------------------------------------------------------------
template <bool STRICT>
struct TexFetch;
template <>
struct TexFetch<false> {
    typedef float4 texel; // c / 255 per channel
    static __device__ __forceinline__ texel fetch(const FrameParams& p, float x, float y, float z) {
        return tex3D<float4>((cudaTextureObject_t)p.tex_unorm, x, y, z);
    }
    static __device__ __forceinline__ texel zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <>
struct TexFetch<true> {
    typedef uchar4 texel; // raw bytes
    static __device__ __forceinline__ texel fetch(const FrameParams& p, float x, float y, float z) {
        return tex3D<uchar4>((cudaTextureObject_t)p.tex_raw, x, y, z);
    }
    static __device__ __forceinline__ texel zero() { return make_uchar4(0, 0, 0, 0); }
};

// The instrumented pass can also record WHICH voxels a frame fetches (xn_render_touch_pass): one bit
// per voxel of the x-major linear order, so a byte of the map is one 32-byte sector of that layout.
template <bool STRICT, bool STATS>
struct TexFetchS : TexFetch<STRICT> {
    typedef typename TexFetch<STRICT>::texel texel;
    static __device__ __forceinline__ texel fetch(const FrameParams& p, float x, float y, float z) {
        if (STATS && p.touch_bits) {
            const int vx = __float2int_rd(x), vy = __float2int_rd(y), vz = __float2int_rd(z);
            if ((uint32_t)vx < p.nx && (uint32_t)vy < p.ny && (uint32_t)vz < p.nz) { // border texels are not memory
                const uint64_t i = (uint64_t)vx + (uint64_t)p.nx * ((uint64_t)vy + (uint64_t)p.ny * (uint64_t)vz);
                atomicOr(p.touch_bits + (i >> 5), 1u << (uint32_t)(i & 31u));
            }
        }
        return TexFetch<STRICT>::fetch(p, x, y, z);
    }
};

// emission accumulator of the texture path: strict = shader order on exact c / 255,
// fast = fma on the texture unit's c / 255 (scaled by the emission coefficient only)
template <bool STRICT>
struct TexAccum;
template <>
struct TexAccum<false> {
    float r = 0.f, g = 0.f, b = 0.f;
    __device__ __forceinline__ void add(float4 c, float len) {
        r = __fmaf_rn(c.x, len, r);
        g = __fmaf_rn(c.y, len, g);
        b = __fmaf_rn(c.z, len, b);
    }
    __device__ __forceinline__ f3 finish(float ec) const { return F3(r * ec, g * ec, b * ec); }
};
template <>
struct TexAccum<true> {
    float r = 0.f, g = 0.f, b = 0.f;
    __device__ __forceinline__ void add(uchar4 c, float len) {
        r += unorm8_exact((float)c.x) * len;
        g += unorm8_exact((float)c.y) * len;
        b += unorm8_exact((float)c.z) * len;
    }
    __device__ __forceinline__ f3 finish(float ec) const { return F3(r * ec, g * ec, b * ec); }
};

template <bool STATS, bool STRICT>
__global__ void __launch_bounds__(BLOCK_THREADS, XN_TEX_MIN_BLOCKS) dda_tex_kernel(const __grid_constant__ FrameParams p) {
    typedef TexFetchS<STRICT, STATS> TF;
    typedef typename TF::texel texel;
    uint32_t ix, iy;
    thread_pixel<true>(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const float side = fmaxf((float)p.nx, fmaxf((float)p.ny, (float)p.nz));
    f3 ro = F3(p.pos[0] * side, p.pos[1] * side, p.pos[2] * side);
    const float ec = voxel_emission_coeff(p, rd) / side;

    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);
    const f3 bmin = F3(-bias.x, -bias.y, -bias.z);
    const f3 bmax = F3((float)p.model_dim[0] * rrd.x - bias.x, (float)p.model_dim[1] * rrd.y - bias.y,
                       (float)p.model_dim[2] * rrd.z - bias.z);
    float t_min = max_elem(F3(gmin(bmin.x, bmax.x), gmin(bmin.y, bmax.y), gmin(bmin.z, bmax.z)));
    const float t_max = min_elem(F3(gmax(bmin.x, bmax.x), gmax(bmin.y, bmax.y), gmax(bmin.z, bmax.z)));

    TexAccum<STRICT> acc;
    if (!(t_min > t_max)) {
        t_min = gmax(t_min, 0.0f);
        ro = F3(ro.x + rd.x * t_min, ro.y + rd.y * t_min, ro.z + rd.z * t_min);
        const float tdx = fabsf(rrd.x), tdy = fabsf(rrd.y), tdz = fabsf(rrd.z);
        const f3 sg = F3(gsign(rd.x), gsign(rd.y), gsign(rd.z));
        float sdx = (sg.x * ((floorf(ro.x) - ro.x) + 0.5f) + 0.5f) * tdx;
        float sdy = (sg.y * ((floorf(ro.y) - ro.y) + 0.5f) + 0.5f) * tdy;
        float sdz = (sg.z * ((floorf(ro.z) - ro.z) + 0.5f) + 0.5f) * tdz;
        // ivec3(ro) truncates; texel centres are exact in binary32 (coordinates below 2^22)
        float fx = (float)(int)ro.x + 0.5f, fy = (float)(int)ro.y + 0.5f, fz = (float)(int)ro.z + 0.5f;

        texel v = TF::fetch(p, fx, fy, fz);
        float t = 0.0f;
        const float t_end = t_max - t_min;
#define XN_TEX_STEP(DT)                                            \
    {                                                              \
        const float t0 = fminf(sdx, fminf(sdy, sdz));              \
        const bool mx = sdx == t0, my = sdy == t0, mz = sdz == t0; \
        DT = t0 - t;                                               \
        t = t0;                                                    \
        if (mx) { sdx += tdx; fx += sg.x; }                        \
        if (my) { sdy += tdy; fy += sg.y; }                        \
        if (mz) { sdz += tdz; fz += sg.z; }                        \
        st.step();                                                 \
        st.read(4);                                                \
    }
        // as in dda_kernel: while t < t_end - 4.5 td_min four more steps are certain, so they run
        // as one trip whose four fetches are issued together and consumed one trip later
        const float t_lim4 = t_end - 4.5f * fminf(tdx, fminf(tdy, tdz));
        if (t < t_lim4) {
            texel a0 = TF::zero(), a1 = TF::zero(), a2 = TF::zero(), a3 = v, b0, b1, b2, b3;
            float ad1 = 0.f, ad2 = 0.f, ad3 = 0.f, bd1, bd2, bd3;
#define XN_TEX_TRIP(N, P)                                  \
    {                                                      \
        float d0;                                          \
        XN_TEX_STEP(d0) N##0 = TF::fetch(p, fx, fy, fz);    \
        XN_TEX_STEP(N##d1) N##1 = TF::fetch(p, fx, fy, fz); \
        XN_TEX_STEP(N##d2) N##2 = TF::fetch(p, fx, fy, fz); \
        XN_TEX_STEP(N##d3) N##3 = TF::fetch(p, fx, fy, fz); \
        acc.add(P##0, P##d1);                              \
        acc.add(P##1, P##d2);                              \
        acc.add(P##2, P##d3);                              \
        acc.add(P##3, d0);                                 \
    }
            for (;;) {
                XN_TEX_TRIP(b, a)
                if (!(t < t_lim4)) {
                    acc.add(b0, bd1);
                    acc.add(b1, bd2);
                    acc.add(b2, bd3);
                    v = b3;
                    break;
                }
                XN_TEX_TRIP(a, b)
                if (!(t < t_lim4)) {
                    acc.add(a0, ad1);
                    acc.add(a1, ad2);
                    acc.add(a2, ad3);
                    v = a3;
                    break;
                }
            }
#undef XN_TEX_TRIP
        }
        while (t < t_end) {
            float dt;
            XN_TEX_STEP(dt)
            const texel vn = TF::fetch(p, fx, fy, fz);
            acc.add(v, dt);
            v = vn;
        }
#undef XN_TEX_STEP
    }
    store_result(p, ix, iy, acc.finish(ec), st);
}

// ---------------------------------------------------------------------------------
// DDA over the texture residency with the skip table (resources/dda.comp:13-73, same arithmetic).
//
// Every ray still takes every step of the reference's march -- min of the side distances, the
// axes that tie step together, t advances by the same binary32 additions -- so the sequence of
// voxels, the segment lengths and the per-ray step counts are those of dda.comp.  What changes is
// the fetch at dda.comp:45: on volumes that are mostly one colour (TNG gas: 90-96 % of voxels) the
// skip table (xn_util_kernels.cu) tells a ray "travelling your way from this brick, every texel
// within R_i voxels on axis i is colour uc", and texels the ray can know are not read.
//
// The promise is kept as a time, not a count.  Axis i takes its k-th step from now in the
// iteration that ends at the side distance after k - 1 more additions of td_i, so the voxel a
// step lands in is inside the promised box iff the step ends before
//     t_safe = min_i(sd_i + R_i td_i) (1 - 2^-14)
// (the factor bounds the rounding of up to ~1000 repeated additions from below).  A step advances
// t by at most td_min, so a trip of four steps that starts before t_safe - 4.5 td_min lands on
// promised texels only: one comparison per trip decides whether its four TEX are issued.  The
// table is consulted again two trips before the promise runs out (usually the ray has entered
// bricks that promise more by then), or at the exit of a mixed brick; a late or early look-up
// only costs fetches, never correctness.
//
// Texels are consumed one trip after they were requested, as in dda_tex_kernel.  In the fast mode
// the length of a known step goes to klen (through a 0/1 factor, on the FMA pipe: the ALU pipe --
// min, compares -- is what bounds this kernel) and uc * klen is added when uc changes or the ray
// ends; the strict mode accumulates uc * dt per step in the shader's order.  While every active
// lane of the warp is in a known trip, the warp runs bare trips: geometry only.
// ---------------------------------------------------------------------------------
#ifndef XN_SKIP_MIN_BLOCKS
#define XN_SKIP_MIN_BLOCKS 4
#endif
// known stretches: 0 = full trips only, 1 = counted bare trips (geometry only, warp-wide count),
// 2 = per-lane closed-form jump over all steps that end before the promise does (dda_axis_jump)
#ifndef XN_SKIP_BARE
#define XN_SKIP_BARE 2
#endif
// jump when a lane of the warp has at least this many steps of known texels ahead
#ifndef XN_SKIP_JUMP_MIN
#define XN_SKIP_JUMP_MIN 6
#endif
// 1 = a jump takes every step that ends before the promise does (the next trip fetches); 0 = it
// stops a trip short of that, and a trip of known steps follows
#ifndef XN_SKIP_FULL
#define XN_SKIP_FULL 1
#endif
// look-up + jump rounds between two trips: after a jump the ray sits in the last promised voxel,
// whose brick usually promises more
// 1 = trips always fetch (known texels included: with the jump taking the known stretches, 0.1 % of
// the steps; the known / fetched bookkeeping of every step costs more than those fetches: cfg4
// 1504 -> 1714); 0 = a trip that starts well inside a promise does not fetch
#ifndef XN_SKIP_TRIP_FETCH
#define XN_SKIP_TRIP_FETCH 1
#endif
#ifndef XN_SKIP_LOOK_TRIPS
#define XN_SKIP_LOOK_TRIPS 4
#endif
#ifndef XN_SKIP_HOPS
#define XN_SKIP_HOPS 1
#endif
// rounds after the first need this many lanes of the warp able to jump (1 = any lane)
#ifndef XN_SKIP_HOP_LANES
#define XN_SKIP_HOP_LANES 1
#endif
// 1 = a jump may start while a fetched texel is pending (its length is kept in plen)
#ifndef XN_SKIP_PLEN
#define XN_SKIP_PLEN 1
#endif
// look the table up this many trips before the promise runs out
#ifndef XN_SKIP_EARLY
#define XN_SKIP_EARLY 2
#endif
// experiment counters written to the STATS pass's byte counts instead of the algorithmic bytes:
// 1 = steps taken in bare trips, 2 = table look-ups, 3 = texels fetched, 4 = known steps in full trips
#ifndef XN_SKIP_DEBUG
#define XN_SKIP_DEBUG 0
#endif
// a table colour (rgb bytes) as the texel type of the mode
template <bool STRICT>
struct TexUniform;
template <>
struct TexUniform<false> {
    static __device__ __forceinline__ float4 texel_of(uint32_t rgb) {
        const float r = 1.0f / 255.0f;
        return make_float4((float)(rgb & 0xFFu) * r, (float)((rgb >> 8) & 0xFFu) * r, (float)((rgb >> 16) & 0xFFu) * r, 0.f);
    }
};
template <>
struct TexUniform<true> {
    static __device__ __forceinline__ uchar4 texel_of(uint32_t rgb) {
        return make_uchar4((unsigned char)(rgb & 0xFFu), (unsigned char)((rgb >> 8) & 0xFFu),
                           (unsigned char)((rgb >> 16) & 0xFFu), 0);
    }
};

// Closed form of the march's repeated additions (XN_SKIP_BARE == 2).  A side distance advances by
//     s <- fl(s + d)       (binary32, round to nearest even, d > 0)
// once per crossing of its axis, whatever the other axes do.  While s stays inside one binade
// [2^e, 2^(e+1)) every sum is rounded to a multiple of the SAME u = 2^(e-23), and s itself is such a
// multiple, so fl(s + d) = s + q with q = d rounded to a multiple of u: the crossings of an axis are
// an exact arithmetic progression inside a binade, s_j = s + j q, with q read off one real addition
// (q = fl(s + d) - s, exact).  [If d ends exactly on u / 2 the rounding is a tie, resolved towards
// the even multiple: the result of any such addition is even, and from an even s the increment is
// again constant -- so the progression is used from an even s only.]  All crossings below a time T
// are therefore taken in O(binades) operations instead of O(crossings), and land on bit-identical
// values: j = the largest count with s + j q < min(T, 2^(e+1)) comes from an under-estimate of the
// quotient plus exact single steps (every s + j q below 2^(e+1) is representable, so fma(j, q, s) and
// the corrections are exact); the addition that leaves the binade, and the one after it, are real ones.
// On return s is the first crossing not below T, kf has grown by the crossings taken and t_last is
// the latest of them.  (Prototype checked against the sequential additions on 3 10^5 random
// (s, d, T) including tie-prone d: tools/jump_proto.py.)
//
// XN_SKIP_ONE_BINADE = 1 caps T at the top of every axis' current binade (dda_binade_top), so that a
// jump never continues in the next binade: a lane crosses a power of two in about 2 % of its jumps
// at large t, but a warp repeats the loop body when any of its lanes has to (42 % of the warps, ncu
// r02g).  Measured slower (cfg4 1818 -> 1750, cfg1 9313 -> 7786): near the start of a ray binades
// are a few steps long and a promise spans several, and a lane cut short asks for another jump round
// (130 instructions for the warp) where the repeated loop body cost 40.  Off.
#ifndef XN_SKIP_ONE_BINADE
#define XN_SKIP_ONE_BINADE 0
#endif
__device__ __forceinline__ float dda_binade_top(float s) { // 2^(e+1) for s in [2^e, 2^(e+1))
    return __uint_as_float((__float_as_uint(s) & 0x7F800000u) + 0x00800000u);
}
__device__ __forceinline__ void dda_axis_jump(float& s, const float d, const float T, float& t_last, float& kf) {
    while (s < T) {
        const float s1 = s + d, q = s1 - s; // one real addition: the next crossing, and the increment
        const uint32_t sb = __float_as_uint(s), eb = sb & 0x7F800000u;
        const float top = __uint_as_float(eb + 0x00800000u);    // 2^(e+1)
        const float half_u = __uint_as_float(eb - (24u << 23)); // 2^(e-24)
        if (s1 < top && (fabsf(d - q) != half_u || (sb & 1u) == 0u)) {
            // s, s + q, ... below hi = min(T, 2^(e+1)) are all crossings taken; the one after the
            // last of them is a real addition again (it may leave the binade, or pass T)
            const float hi = XN_SKIP_ONE_BINADE ? T : fminf(T, top);
            float jf = floorf(__fdividef(hi - s, q) * 0.99999f);
            float sj = __fmaf_rn(jf, q, s);
            while (sj + q < hi) {
                sj += q;
                jf += 1.0f;
            }
            t_last = fmaxf(t_last, sj);
            s = sj + d;
            kf += jf + 1.0f;
        } else {
            t_last = fmaxf(t_last, s);
            s = s1;
            kf += 1.0f;
        }
    }
}

template <bool STATS, bool STRICT>
__global__ void __launch_bounds__(BLOCK_THREADS, XN_SKIP_MIN_BLOCKS)
    dda_skip_tex_kernel(const __grid_constant__ FrameParams p) {
    typedef TexFetchS<STRICT, STATS> TF;
    typedef typename TF::texel texel;
    uint32_t ix, iy;
    thread_pixel<true>(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const float side = fmaxf((float)p.nx, fmaxf((float)p.ny, (float)p.nz));
    f3 ro = F3(p.pos[0] * side, p.pos[1] * side, p.pos[2] * side);
    const float ec = voxel_emission_coeff(p, rd) / side;

    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);
    const f3 bmin = F3(-bias.x, -bias.y, -bias.z);
    const f3 bmax = F3((float)p.model_dim[0] * rrd.x - bias.x, (float)p.model_dim[1] * rrd.y - bias.y,
                       (float)p.model_dim[2] * rrd.z - bias.z);
    float t_min = max_elem(F3(gmin(bmin.x, bmax.x), gmin(bmin.y, bmax.y), gmin(bmin.z, bmax.z)));
    const float t_max = min_elem(F3(gmax(bmin.x, bmax.x), gmax(bmin.y, bmax.y), gmax(bmin.z, bmax.z)));

    TexAccum<STRICT> acc;
    uint32_t uc = 0;   // colour of the current promise
    float klen = 0.0f; // fast mode: length of known steps not yet added as uc * klen
    float plen = 0.0f; // fast mode: length a jump gave the fetched texel still waiting in the pipeline
    if (!(t_min > t_max)) {
        t_min = gmax(t_min, 0.0f);
        ro = F3(ro.x + rd.x * t_min, ro.y + rd.y * t_min, ro.z + rd.z * t_min);
        const float tdx = fabsf(rrd.x), tdy = fabsf(rrd.y), tdz = fabsf(rrd.z);
        const f3 sg = F3(gsign(rd.x), gsign(rd.y), gsign(rd.z));
        float sdx = (sg.x * ((floorf(ro.x) - ro.x) + 0.5f) + 0.5f) * tdx;
        float sdy = (sg.y * ((floorf(ro.y) - ro.y) + 0.5f) + 0.5f) * tdy;
        float sdz = (sg.z * ((floorf(ro.z) - ro.z) + 0.5f) + 0.5f) * tdz;
        // ivec3(ro) truncates; texel centres are exact in binary32 (coordinates below 2^22)
        float fx = (float)(int)ro.x + 0.5f, fy = (float)(int)ro.y + 0.5f, fz = (float)(int)ro.z + 0.5f;

        const uint32_t S = p.skip_shift, BM = (1u << S) - 1u;
        // voxels left to the brick face in the direction of travel: c ^ BM going up, c going down
        const uint32_t xm = sg.x > 0.0f ? BM : 0u, ym = sg.y > 0.0f ? BM : 0u, zm = sg.z > 0.0f ? BM : 0u;
        // octant of travel -> which radius byte of a table entry applies to this ray
        const uint32_t oct = (sg.x > 0.0f ? 1u : 0u) | (sg.y > 0.0f ? 2u : 0u) | (sg.z > 0.0f ? 4u : 0u);
        const uint32_t osh = 8u * (oct & 3u);
        const float td_min = fminf(tdx, fminf(tdy, tdz));
        const float td45 = 4.5f * td_min;
        const float td_look = (float)(4 * XN_SKIP_LOOK_TRIPS) * td_min; // spacing floor of table look-ups, in trips
        const float inv_trip = 0.999999f / (4.00390625f * td_min); // trips per unit of t, rounded down a little
        const float itdx = 1.0f / tdx, itdy = 1.0f / tdy, itdz = 1.0f / tdz;
        const float isum = (itdx + itdy) + itdz; // steps per unit of t
        // longest bare run, in trips minus one: an axis' side distance takes at most 4 n + 1 additions in
        // n trips, each rounded by at most 2^-24 (t_end + td_i), so the step count recovered from it is
        // off by less than (4 n + 1) 2^-24 (t_end / td_min + 1) -- kept below a quarter
        const uint32_t n_cap = (uint32_t)fminf(1023.0f, fmaxf(65536.0f * td_min / ((t_max - fmaxf(t_min, 0.0f)) + td_min) * 15.9f - 1.0f, 0.0f));
        float t_safe = -1.0f; // a step that ends before t_safe lands on a texel of colour uc
        float t_look = 0.0f;  // consult the table at the first trip boundary at or after this time
        texel uct = TF::zero(); // uc as a texel (strict mode)

        texel v = TF::fetch(p, fx, fy, fz);
        // 1.0 while the texel pending its segment was fetched, 0.0 while it is a known one (colour uc)
        float pf = 1.0f;
        float t = 0.0f;
        const float t_end = t_max - t_min;
        const float t_lim4 = t_end - td45;

        // geometry of one step of dda.comp:41-50; DTV = its length
#define XN_SKIP_GEOM(DTV)                                          \
    {                                                              \
        const float t0 = fminf(sdx, fminf(sdy, sdz));              \
        const bool mx = sdx == t0, my = sdy == t0, mz = sdz == t0; \
        DTV = t0 - t;                                              \
        t = t0;                                                    \
        if (mx) { sdx += tdx; fx += sg.x; }                        \
        if (my) { sdy += tdy; fy += sg.y; }                        \
        if (mz) { sdz += tdz; fz += sg.z; }                        \
    }
        // One step inside a trip.  The step's length belongs to the texel requested by the PREVIOUS
        // step: FP = 1.0 if that one was fetched (DT = length) or 0.0 if it is known (length -> klen).
        // FETCH (constant over the trip): request the texel of the voxel stepped into.  PLAIN: the
        // previous texel is certainly a fetched one (no known-length bookkeeping).
#define XN_SKIP_STEP(DT, FP, TEXEL, FETCH, PLAIN)                  \
    {                                                              \
        float dt_;                                                 \
        XN_SKIP_GEOM(dt_)                                          \
        st.step();                                                 \
        if (!XN_SKIP_DEBUG) st.read(4);                            \
        if (STRICT || PLAIN) {                                     \
            DT = dt_;                                              \
        } else {                                                   \
            DT = dt_ * FP;                                         \
            klen = __fmaf_rn(dt_, 1.0f - FP, klen);                \
        }                                                          \
        if (XN_SKIP_DEBUG == 3 && FETCH) st.read(1);               \
        if (XN_SKIP_DEBUG == 4 && !FETCH) st.read(1);              \
        if (FETCH) TEXEL = TF::fetch(p, fx, fy, fz);               \
        else if (STRICT) TEXEL = uct;                              \
    }
        // Table look-up at the current voxel (every active lane of the warp does it when any lane's
        // promise is about to run out: a fresh promise never hurts, and the lanes' next look-ups
        // move away together).  PEND = the texel pending its segment: if the promise changes colour
        // while that texel is a known one, it is made explicit first.
#define XN_SKIP_LOOKUP(PEND)                                                                                       \
    {                                                                                                              \
        if (XN_SKIP_DEBUG == 2) st.read(1);                                                                        \
        const int vx = __float2int_rd(fx), vy = __float2int_rd(fy), vz = __float2int_rd(fz);                        \
        /* beyond the table's border layer everything is border colour: clamp onto the layer */                    \
        const uint32_t bx = (uint32_t)min(max((vx >> S) + 1, 0), (int)p.skip_dim[0] - 1);                           \
        const uint32_t by = (uint32_t)min(max((vy >> S) + 1, 0), (int)p.skip_dim[1] - 1);                           \
        const uint32_t bz = (uint32_t)min(max((vz >> S) + 1, 0), (int)p.skip_dim[2] - 1);                           \
        const uint4 e = __ldg(p.skip_table + ((bz * p.skip_dim[1] + by) * p.skip_dim[0] + bx));                     \
        const uint32_t k = ((oct & 4u ? e.w : e.z) >> osh) & 0xFFu; /* bricks of this colour ahead, 0 = mixed */    \
        const uint32_t ext = k > 1u ? (k - 1u) << S : 0u;                                                          \
        const float rx = (float)((((uint32_t)vx ^ xm) & BM) + ext), ry = (float)((((uint32_t)vy ^ ym) & BM) + ext), \
                    rz = (float)((((uint32_t)vz ^ zm) & BM) + ext);                                                 \
        const float te = fminf(sdx + rx * tdx, fminf(sdy + ry * tdy, sdz + rz * tdz)) * 0.99993896484375f;          \
        float tl = te; /* mixed brick, or a short promise: look again when the ray is past it */                   \
        if (k != 0u) {                                                                                             \
            const uint32_t col = e.x & 0x00FFFFFFu;                                                                \
            if (col != uc) {                                                                                       \
                if (!STRICT) {                                                                                     \
                    acc.add(TexUniform<STRICT>::texel_of(uc), klen);                                               \
                    klen = 0.0f;                                                                                   \
                    if (pf == 0.0f) {                                                                              \
                        if (plen != 0.0f) { /* PEND still holds a fetched texel that a jump closed */              \
                            acc.add(PEND, plen);                                                                   \
                            plen = 0.0f;                                                                           \
                        }                                                                                          \
                        PEND = TexUniform<STRICT>::texel_of(uc);                                                   \
                        pf = 1.0f;                                                                                 \
                    }                                                                                              \
                }                                                                                                  \
                uc = col;                                                                                          \
                if (STRICT) uct = TexUniform<STRICT>::texel_of(col);                                               \
                t_safe = te;                                                                                       \
            } else {                                                                                               \
                t_safe = fmaxf(t_safe, te); /* both promises hold */                                               \
            }                                                                                                      \
            /* a promise worth at least three trips: renew it XN_SKIP_EARLY trips before it runs out */           \
            if (t_safe - t > 3.0f * td45) tl = t_safe - (float)XN_SKIP_EARLY * td45;                              \
        }                                                                                                          \
        t_look = fmaxf(tl, t + td_look); /* and never within the next two trips */                                 \
    }
        // Between full trips: look-up when due, then as many bare trips -- four steps of geometry
        // and nothing else -- as EVERY active lane of the warp is certain to spend on promised
        // texels.  A step advances t by at most td_min, so lane i has at least
        // floor((min(t_safe, t_end) - 4.5 td_min - t) / (4 td_min)) such trips ahead (none while a
        // fetched texel is pending); the warp takes the minimum (one REDUX) and runs a counted loop.
#define XN_SKIP_BARE_TRIPS(PEND)                                                                    \
    {                                                                                               \
        const unsigned am = __activemask();                                                         \
        if (__any_sync(am, !(t < t_look))) XN_SKIP_LOOKUP(PEND)                                      \
        if (!STRICT && XN_SKIP_BARE == 2) {                                                         \
          _Pragma("unroll 1") for (int hop = 0; hop < XN_SKIP_HOPS; ++hop) {                         \
            if (hop != 0 && __any_sync(am, !(t < t_look))) XN_SKIP_LOOKUP(PEND)                      \
            /* every step that ends before T lands on a promised texel and leaves a whole trip before */ \
            /* the end of the ray: take them all at once, each lane its own T                        */ \
            float T = fminf(XN_SKIP_FULL ? t_safe : t_safe - td45, t_lim4);                          \
            if (XN_SKIP_ONE_BINADE)                                                                 \
                T = fminf(T, fminf(dda_binade_top(sdx), fminf(dda_binade_top(sdy), dda_binade_top(sdz)))); \
            /* a lane jumps if at least one step ends before T.  If the texel pending its segment was */ \
            /* fetched (the trip before this fetched), the first of those steps closes it: its       */ \
            /* length goes to plen, which the next trip adds when it consumes that texel -- so a     */ \
            /* ray goes from fetching to jumping without a trip of known steps in between            */ \
            const float tfirst = fminf(sdx, fminf(sdy, sdz));                                       \
            const bool can = tfirst < T && (XN_SKIP_PLEN || pf == 0.0f);                            \
            /* the first round jumps if any lane can; a further round before the next trip only if */ \
            /* at least XN_SKIP_HOP_LANES lanes can (the others would wait for them)               */ \
            const unsigned cm = __ballot_sync(am, can && (T - t) * isum >= (float)XN_SKIP_JUMP_MIN);  \
            if (hop == 0 ? cm == 0u : __popc(cm) < XN_SKIP_HOP_LANES) break;                        \
            {                                                                                       \
                if (can) {                                                                          \
                    float tl = t, kx = 0.0f, ky = 0.0f, kz = 0.0f;                                  \
                    float tb = t;                                                                   \
                    if (XN_SKIP_PLEN && pf != 0.0f) {                                               \
                        plen = tfirst - t;                                                          \
                        tb = tfirst;                                                                \
                        pf = 0.0f;                                                                  \
                    }                                                                               \
                    if (STATS) {                                                                    \
                        /* the instrumented build takes the same steps one by one (counting them) */ \
                        /* and checks the closed form against them: a mismatch poisons the count  */ \
                        float cx = sdx, cy = sdy, cz = sdz;                                         \
                        dda_axis_jump(cx, tdx, T, tl, kx);                                          \
                        dda_axis_jump(cy, tdy, T, tl, ky);                                          \
                        dda_axis_jump(cz, tdz, T, tl, kz);                                          \
                        float qx = 0.0f, qy = 0.0f, qz = 0.0f, tq = t;                              \
                        while (fminf(sdx, fminf(sdy, sdz)) < T) {                                   \
                            const float t0 = fminf(sdx, fminf(sdy, sdz));                           \
                            const bool mx = sdx == t0, my = sdy == t0, mz = sdz == t0;              \
                            tq = t0;                                                                \
                            if (mx) { sdx += tdx; qx += 1.0f; }                                     \
                            if (my) { sdy += tdy; qy += 1.0f; }                                     \
                            if (mz) { sdz += tdz; qz += 1.0f; }                                     \
                            st.step();                                                              \
                            st.read(XN_SKIP_DEBUG == 0 ? 4u : (XN_SKIP_DEBUG == 1 ? 1u : 0u));      \
                        }                                                                           \
                        if (cx != sdx || cy != sdy || cz != sdz || kx != qx || ky != qy || kz != qz || tl != tq) \
                            st.steps |= 0x40000000u;                                                \
                    } else {                                                                        \
                        dda_axis_jump(sdx, tdx, T, tl, kx);                                         \
                        dda_axis_jump(sdy, tdy, T, tl, ky);                                         \
                        dda_axis_jump(sdz, tdz, T, tl, kz);                                         \
                    }                                                                               \
                    fx = __fmaf_rn(sg.x, kx, fx);                                                   \
                    fy = __fmaf_rn(sg.y, ky, fy);                                                   \
                    fz = __fmaf_rn(sg.z, kz, fz);                                                   \
                    klen += tl - tb;                                                                \
                    t = tl;                                                                         \
                }                                                                                   \
            }                                                                                       \
          }                                                                                         \
        } else if (!STRICT && XN_SKIP_BARE == 1) {                                                  \
            /* trips certain to be known: the k-th starts before t + 4 k td_min (1 + 2^-10), and must  */ \
            /* start before min(t_safe, t_end) - 4.5 td_min; +1 because trip 0 starts at t itself; a */ \
            /* negative count converts to 0                                                          */ \
            const float x = __fmaf_rn(fminf(t_safe - td45, t_lim4) - t, inv_trip, 1.0f);              \
            const uint32_t ni = pf == 0.0f ? min(__float2uint_rz(x), n_cap) : 0u;                     \
            const uint32_t n = __reduce_min_sync(am, ni);                                           \
            if (n != 0u) {                                                                          \
                /* positions are not needed while nothing is fetched: the run advances the side */  \
                /* distances only, and the texel centres follow from how often each axis stepped, */ \
                /* k_i = round((sd_i - sd_i before) / td_i) -- exact, the run is capped (n_cap) so  */ \
                /* that the rounding of its additions stays below a quarter of a step             */ \
                const float tb = t, sx0 = sdx, sy0 = sdy, sz0 = sdz;                                \
                for (uint32_t k = 0; k < n; ++k) {                                                  \
                    _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                 \
                        const float t0 = fminf(sdx, fminf(sdy, sdz));                               \
                        const bool mx = sdx == t0, my = sdy == t0, mz = sdz == t0;                  \
                        t = t0;                                                                     \
                        if (mx) sdx += tdx;                                                         \
                        if (my) sdy += tdy;                                                         \
                        if (mz) sdz += tdz;                                                         \
                    }                                                                               \
                }                                                                                   \
                fx = __fmaf_rn(sg.x, rintf((sdx - sx0) * itdx), fx);                                \
                fy = __fmaf_rn(sg.y, rintf((sdy - sy0) * itdy), fy);                                \
                fz = __fmaf_rn(sg.z, rintf((sdz - sz0) * itdz), fz);                                \
                klen += t - tb;                                                                     \
                if (STATS) {                                                                        \
                    st.steps += 4u * n;                                                             \
                    st.read(XN_SKIP_DEBUG == 0 ? 16u * n : (XN_SKIP_DEBUG == 1 ? 4u * n : 0u));     \
                }                                                                                   \
            }                                                                                       \
        }                                                                                           \
    }
#define XN_SKIP_TRIP(N, P)                                  \
    {                                                       \
        const bool fetch = XN_SKIP_TRIP_FETCH || !(t < t_safe - td45); \
        const float nf = fetch ? 1.0f : 0.0f;               \
        float d0;                                           \
        XN_SKIP_STEP(d0, pf, N##0, fetch, false)            \
        XN_SKIP_STEP(N##d1, nf, N##1, fetch, XN_SKIP_TRIP_FETCH) \
        XN_SKIP_STEP(N##d2, nf, N##2, fetch, XN_SKIP_TRIP_FETCH) \
        XN_SKIP_STEP(N##d3, nf, N##3, fetch, XN_SKIP_TRIP_FETCH) \
        pf = nf;                                            \
        acc.add(P##0, P##d1);                               \
        acc.add(P##1, P##d2);                               \
        acc.add(P##2, P##d3);                               \
        acc.add(P##3, STRICT ? d0 : d0 + plen);             \
        plen = 0.0f;                                        \
    }
#define XN_SKIP_FLUSH(P)       \
    {                          \
        acc.add(P##0, P##d1);  \
        acc.add(P##1, P##d2);  \
        acc.add(P##2, P##d3);  \
        v = P##3;              \
    }
        if (t < t_lim4) {
            texel a0 = TF::zero(), a1 = TF::zero(), a2 = TF::zero(), a3 = v;
            texel b0 = TF::zero(), b1 = TF::zero(), b2 = TF::zero(), b3 = TF::zero();
            float ad1 = 0.f, ad2 = 0.f, ad3 = 0.f, bd1 = 0.f, bd2 = 0.f, bd3 = 0.f;
            for (;;) {
                XN_SKIP_BARE_TRIPS(a3)
                if (!(t < t_lim4)) { XN_SKIP_FLUSH(a) break; }
                XN_SKIP_TRIP(b, a)
                if (!(t < t_lim4)) { XN_SKIP_FLUSH(b) break; }
                XN_SKIP_BARE_TRIPS(b3)
                if (!(t < t_lim4)) { XN_SKIP_FLUSH(b) break; }
                XN_SKIP_TRIP(a, b)
                if (!(t < t_lim4)) { XN_SKIP_FLUSH(a) break; }
            }
        }
        // the last few steps, one at a time (the promise in force still applies)
        if (!STRICT && plen != 0.0f) {
            acc.add(v, plen);
            plen = 0.0f;
        }
        while (t < t_end) {
            float dt;
            texel vn = v;
            float dt_;
            XN_SKIP_GEOM(dt_)
            st.step();
            if (!XN_SKIP_DEBUG) st.read(4);
            const bool fetch = !(t < t_safe);
            if (STRICT) {
                dt = dt_;
            } else {
                dt = dt_ * pf;
                klen = __fmaf_rn(dt_, 1.0f - pf, klen);
            }
            if (fetch) vn = TF::fetch(p, fx, fy, fz);
            else if (STRICT) vn = uct;
            acc.add(v, dt);
            v = vn;
            pf = fetch ? 1.0f : 0.0f;
        }
#undef XN_SKIP_FLUSH
#undef XN_SKIP_TRIP
#undef XN_SKIP_BARE_TRIPS
#undef XN_SKIP_LOOKUP
#undef XN_SKIP_STEP
#undef XN_SKIP_GEOM
        if (!STRICT) acc.add(TexUniform<STRICT>::texel_of(uc), klen);
    }
    store_result(p, ix, iy, acc.finish(ec), st);
}

// ---------------------------------------------------------------------------------
// shared octree helpers
// ---------------------------------------------------------------------------------
// Traversal stacks: [level][thread] in shared memory, addressed in the shared window directly
// (one LEA + one LDS/STS per access; generic pointers cost five address instructions here).
__device__ __forceinline__ uint32_t stack_base(const void* smem) {
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * (uint32_t)sizeof(uint2);
    // opaque copy: keeps the address in a register instead of being re-derived (S2R + LEA chain)
    // at every push and pop
    asm volatile("mov.u32 %0, %0;" : "+r"(base));
    return base;
}
__device__ __forceinline__ void stack_store(uint32_t base, uint32_t level, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(base + level * (uint32_t)(BLOCK_THREADS * sizeof(uint2))),
                 "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ uint2 stack_load(uint32_t base, uint32_t level) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "r"(base + level * (uint32_t)(BLOCK_THREADS * sizeof(uint2)))
                 : "memory");
    return r;
}
__device__ __forceinline__ uint2 load_slot(const DNode* __restrict__ nodes, uint32_t node, uint32_t child) {
    return __ldg(&nodes[node].slot[child]);
}
// compact residency: one word per child (leaf: meta bits, internal: compact child index)
// `node` is the WORD offset of the record (compact index * 8, as stored in the parent's word), so the
// address of a child word is one OR and one scaled add
__device__ __forceinline__ uint32_t load_word(const CNode* __restrict__ nodes, uint32_t node, uint32_t child) {
    return __ldg(reinterpret_cast<const uint32_t*>(nodes) + (node | child));
}
__device__ __forceinline__ bool word_is_leaf(uint32_t w) { return (int32_t)w < 0; }
// child descriptor as (index, meta) from either residency
__device__ __forceinline__ uint2 load_child(const DNode* __restrict__ nodes, uint32_t node, uint32_t child) {
    return load_slot(nodes, node, child);
}
__device__ __forceinline__ uint2 load_child(const CNode* __restrict__ nodes, uint32_t node, uint32_t child) {
    const uint32_t w = load_word(nodes, node, child);
    return make_uint2(w, w); // leaf: .y carries flag + colour; internal: .x is the index, .y has bit 31 clear
}

// slab test of the unit cube [0,1]^3 (svo_naive.comp:30-45, svo_rope.comp:72-86)
__device__ __forceinline__ bool unit_cube_slab(f3 rrd, f3 bias, float& t_min, float& t_max) {
    const f3 bmin = F3(-bias.x, -bias.y, -bias.z);
    const f3 bmax = F3(rrd.x - bias.x, rrd.y - bias.y, rrd.z - bias.z);
    t_min = max_elem(F3(gmin(bmin.x, bmax.x), gmin(bmin.y, bmax.y), gmin(bmin.z, bmax.z)));
    t_max = min_elem(F3(gmax(bmin.x, bmax.x), gmax(bmin.y, bmax.y), gmax(bmin.z, bmax.z)));
    if (t_min > t_max) return false;
    t_min = gmax(t_min, 0.0f);
    return true;
}

// descend from (node, meta) at `offset`/`extent` to the leaf containing pos
// (loop body of find(), svo_naive.comp:14-26 == svo_rope.comp:14-26 == svo_rope.comp:33-47)
template <bool STATS, class NODE>
__device__ __forceinline__ void descend(const NODE* __restrict__ nodes, f3 pos, uint32_t& node, uint32_t& meta,
                                        f3& offset, float& extent, RayStats<STATS>& st) {
    for (;;) {
        st.read(4); // is_leaf_depth
        if (meta_is_leaf(meta)) return;
        extent *= 0.5f;
#if XN_DESCEND_FLAGS
        // comparison results as 1.0 / 0.0 (FSET), then exact FMAs: offset += mask * extent is what
        // the shader writes (svo_naive.comp:21-23), and the child index is 4 mx + 2 my + mz
        uint32_t child;
        {
            const float cx = offset.x + extent, cy = offset.y + extent, cz = offset.z + extent;
            asm("{\n\t"
                ".reg .f32 fx, fy, fz, mf;\n\t"
                "set.ge.f32.f32 fx, %4, %7;\n\t"
                "set.ge.f32.f32 fy, %5, %8;\n\t"
                "set.ge.f32.f32 fz, %6, %9;\n\t"
                "fma.rn.f32 %1, fx, %10, %1;\n\t"
                "fma.rn.f32 %2, fy, %10, %2;\n\t"
                "fma.rn.f32 %3, fz, %10, %3;\n\t"
                "fma.rn.f32 mf, fx, 0f40800000, fz;\n\t"
                "fma.rn.f32 mf, fy, 0f40000000, mf;\n\t"
                "cvt.rzi.u32.f32 %0, mf;\n\t"
                "}"
                : "=r"(child), "+f"(offset.x), "+f"(offset.y), "+f"(offset.z)
                : "f"(pos.x), "f"(pos.y), "f"(pos.z), "f"(cx), "f"(cy), "f"(cz), "f"(extent));
        }
#else
        const bool mx = pos.x >= offset.x + extent;
        const bool my = pos.y >= offset.y + extent;
        const bool mz = pos.z >= offset.z + extent;
        const uint32_t child = (mx ? 4u : 0u) + (my ? 2u : 0u) + (mz ? 1u : 0u);
        // offset += vec3(mask) * extent: adding 1.0 * extent or +0.0
        if (mx) offset.x += extent;
        if (my) offset.y += extent;
        if (mz) offset.z += extent;
#endif
        st.read(4); // children[child]
        const uint2 s = load_child(nodes, node, child);
        node = s.x;
        meta = s.y;
    }
}

// chord of the ray through the node box [offset, offset+side] (svo_naive.comp:56-60)
__device__ __forceinline__ void node_slab(f3 offset, float side, f3 rrd, f3 bias, float& u_min, float& u_max,
                                          f3& far) {
    const f3 nmin = F3(offset.x * rrd.x - bias.x, offset.y * rrd.y - bias.y, offset.z * rrd.z - bias.z);
    const f3 nmax = F3((offset.x + side) * rrd.x - bias.x, (offset.y + side) * rrd.y - bias.y,
                       (offset.z + side) * rrd.z - bias.z);
#if XN_SLAB_FMNMX
    // FMNMX instead of the compare-and-select form of GLSL min / max: they differ only in the sign
    // of a zero result (no NaNs: rrd, offsets and bias are finite), and a zero of either sign gives
    // the same comparisons and the same chord; 8 instructions instead of 26
    far = F3(fmaxf(nmin.x, nmax.x), fmaxf(nmin.y, nmax.y), fmaxf(nmin.z, nmax.z));
    u_min = fmaxf(fminf(nmin.x, nmax.x), fmaxf(fminf(nmin.y, nmax.y), fminf(nmin.z, nmax.z)));
    u_max = fminf(far.x, fminf(far.y, far.z));
#else
    far = F3(gmax(nmin.x, nmax.x), gmax(nmin.y, nmax.y), gmax(nmin.z, nmax.z));
    u_min = max_elem(F3(gmin(nmin.x, nmax.x), gmin(nmin.y, nmax.y), gmin(nmin.z, nmax.z)));
    u_max = min_elem(far);
#endif
}

// ---------------------------------------------------------------------------------
// svo_naive (resources/svo_naive.comp:29-89)
// ---------------------------------------------------------------------------------
template <bool STATS, bool STRICT>
__global__ void __launch_bounds__(BLOCK_THREADS, XN_SVO_MIN_BLOCKS) svo_naive_kernel(const __grid_constant__ FrameParams p) {
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;
    const float MIN_STEP_SIZE = 0.00001f;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const f3 ro = F3(p.pos[0], p.pos[1], p.pos[2]);
    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);

    Accum<STRICT> acc;
    float t_min, t_max;
    if (unit_cube_slab(rrd, bias, t_min, t_max)) {
        float t = t_min + MIN_STEP_SIZE;
        while (t < t_max) {
            const f3 pt = F3(t * rd.x + ro.x, t * rd.y + ro.y, t * rd.z + ro.z);
            uint32_t node = 0, meta = p.root_meta;
            f3 offset = F3(0.f, 0.f, 0.f);
            float side = 1.0f;
#if XN_NAIVE_TOP_TABLE
            {
                // The shader's find() starts at the root (svo_naive.comp:50-52).  Each of its decisions,
                // `pos >= offset + extent`, compares against an odd multiple of the child size, so the
                // child taken at level l is bit l of the coordinate's binary expansion -- for any
                // binary32 pos: pos < 0 fails every comparison (bits 0), pos >= 1 passes every one
                // (bits 1).  The first K = min(depth, 8) levels are therefore a function of the cell
                // floor(pos * 2^K), looked up in a table built at upload: one load replaces K
                // dependent ones, and offset = cell * 2^-depth is the same exact sum of powers of
                // two the shader accumulates.  The instrumented build counts the reads of those levels.
                const uint32_t K = p.top_levels;
                const int cells = 1 << K;
                const float fcells = (float)cells;
                const int cx = min(max(__float2int_rd(pt.x * fcells), 0), cells - 1);
                const int cy = min(max(__float2int_rd(pt.y * fcells), 0), cells - 1);
                const int cz = min(max(__float2int_rd(pt.z * fcells), 0), cells - 1);
                const uint32_t w = __ldg(p.top_table + (((uint32_t)cx << (2u * K)) | ((uint32_t)cy << K) | (uint32_t)cz));
                const uint32_t d = word_is_leaf(w) ? meta_depth(w) : K; // levels taken
                const uint32_t drop = K - d;
                side = __int_as_float((127 - (int)d) << 23); // exp2(-d)
                offset = F3((float)(cx >> drop) * side, (float)(cy >> drop) * side, (float)(cz >> drop) * side);
                node = w;
                meta = word_is_leaf(w) ? w : 0u;
                st.read(8u * d); // is_leaf_depth + children[c] of the levels above
            }
#endif
            descend(p.cnodes, pt, node, meta, offset, side, st);

            float u_min, u_max;
            f3 far;
            node_slab(offset, side, rrd, bias, u_min, u_max, far);
            // fmaxf: same values as GLSL max here (finite operands; a zero's sign cannot reach the image)
            u_min = fmaxf(u_min, 0.0f);
            const float step = fmaxf(u_max - u_min, MIN_STEP_SIZE);
            t += step;

            st.read(4); // color
            acc.add(meta, step);
            st.step();
        }
    }
    store_result(p, ix, iy, acc.finish(voxel_emission_coeff(p, rd)), st);
}

// ---------------------------------------------------------------------------------
// svo_df (resources/svo_df.comp:6-85): exhaustive depth-first visit, children in index order.
//
// The shader slab-tests the eight children of a node one loop iteration at a time, rebuilding
// each child's corner from the previous one with three mod()s.  Here the eight tests of a node
// are done together when the node is entered: on every axis the children's boxes are bounded by
// three planes (corner, corner + side, corner + 2 side), whose ray parameters are computed once
// with exactly the shader's operations -- pos * rrd - bias and (pos + side) * rrd - bias, where
// pos is the corner or corner + side -- so every child's (t_min, t_max) is the value the shader
// computes, the hit mask is the shader's `t_min < t_max && t_max > 0`, and only hit children are
// visited, in index order (the accumulation order of the shader).  A stack entry keeps the
// children still to visit and is written only when there are any; the corner is restored on a pop
// by rounding down to the node's size (exact: corners are sums of powers of two).  The instrumented build counts the iterations and reads of
// the shader's loop: 8 iterations per entered node, a read per hit, per leaf, per pop.
// ---------------------------------------------------------------------------------
struct DfPlanes {
    // per axis: min / max ray parameter of the lower (0) and upper (1) half of the node
    float lo0x, hi0x, lo1x, hi1x, lo0y, hi0y, lo1y, hi1y, lo0z, hi0z, lo1z, hi1z;
};
__device__ __forceinline__ void df_axis(float P, float side, float rrd, float bias, float& lo0, float& hi0, float& lo1,
                                        float& hi1) {
    const float pb = P + side;              // corner of the upper half = pos + side of the lower half
    const float a = P * rrd - bias;         // bmin of the lower half
    const float b = pb * rrd - bias;        // bmax of the lower half = bmin of the upper half
    const float c = (pb + side) * rrd - bias; // bmax of the upper half
    // FMNMX instead of the compare-and-select form of GLSL min / max: the two differ only in the
    // sign of a zero result (no NaNs here: rrd and the corners are finite), and a zero of either sign
    // gives the same hit decision and the same chord
    lo0 = fminf(a, b), hi0 = fmaxf(a, b);
    lo1 = fminf(b, c), hi1 = fmaxf(b, c);
}
__device__ __forceinline__ void df_planes(f3 P, float side, f3 rrd, f3 bias, DfPlanes& q) {
    df_axis(P.x, side, rrd.x, bias.x, q.lo0x, q.hi0x, q.lo1x, q.hi1x);
    df_axis(P.y, side, rrd.y, bias.y, q.lo0y, q.hi0y, q.lo1y, q.hi1y);
    df_axis(P.z, side, rrd.z, bias.z, q.lo0z, q.hi0z, q.lo1z, q.hi1z);
}
// (t_min, t_max) of child c (x = bit 2, y = bit 1, z = bit 0), svo_df.comp:27-31
__device__ __forceinline__ void df_child(const DfPlanes& q, uint32_t c, float& t_min, float& t_max) {
    const bool bx = (c & 4u) != 0u, by = (c & 2u) != 0u, bz = (c & 1u) != 0u;
    t_min = fmaxf(bx ? q.lo1x : q.lo0x, fmaxf(by ? q.lo1y : q.lo0y, bz ? q.lo1z : q.lo0z));
    t_max = fminf(bx ? q.hi1x : q.hi0x, fminf(by ? q.hi1y : q.hi0y, bz ? q.hi1z : q.hi0z));
}
__device__ __forceinline__ uint32_t df_hit_mask(const DfPlanes& q) {
#if XN_DF_FLAG_MASK
    // hit bits accumulated as a float on the FMA pipe: each test is two 1.0 / 0.0 comparison flags,
    // their product is the hit, and mask += hit * 2^c is exact (the kernel is ALU-pipe bound)
    float maskf = 0.0f;
#pragma unroll
    for (uint32_t c = 0; c < 8u; ++c) {
        float t_min, t_max, f1, f2;
        df_child(q, c, t_min, t_max); // c is a compile-time constant here: the selects fold away
        asm("set.lt.f32.f32 %0, %2, %3;\n\tset.gt.f32.f32 %1, %3, 0f00000000;" : "=f"(f1), "=f"(f2) : "f"(t_min), "f"(t_max));
        maskf = __fmaf_rn(f1 * f2, (float)(1u << c), maskf);
    }
    return (uint32_t)maskf;
#else
    uint32_t mask = 0;
#pragma unroll
    for (uint32_t c = 0; c < 8u; ++c) {
        float t_min, t_max;
        df_child(q, c, t_min, t_max); // c is a compile-time constant here: the selects fold away
        if (t_min < t_max && t_max > 0.0f) mask |= 1u << c;
    }
    return mask;
#endif
}

// Loop shape.  A node's hit children are consumed in two phases: a LEAF RUN (fetch the child's word;
// a leaf adds colour * chord, ~20 instructions) that stops at the first internal child, and a NODE
// STEP (enter that child, or return to the nearest ancestor with children left: corner, the twelve
// plane parameters and -- when entering -- the eight slab tests, ~110 instructions).  Keeping the
// two apart matters twice: a leaf child no longer pays (predicated off) for the plane arithmetic,
// and the lanes of a warp reconverge after their leaf runs, so the expensive node step runs with
// most of the warp instead of the few lanes that happened to need it in a given iteration
// (ncu r01d: one 143-instruction body per child at 13.5 lanes, pop path at 6.3 lanes).
#ifndef XN_DF_TWO_PHASE
#define XN_DF_TWO_PHASE 1
#endif
#ifndef XN_DF_INLINE
#define XN_DF_INLINE 0
#endif
#ifndef XN_DF_NODE_AT_A_TIME
#define XN_DF_NODE_AT_A_TIME 1
#endif
template <bool STATS, bool STRICT, int LEVELS>
__global__ void __launch_bounds__(BLOCK_THREADS, XN_SVO_MIN_BLOCKS) svo_df_kernel(const __grid_constant__ FrameParams p) {
    __shared__ uint2 stack_mem[LEVELS * BLOCK_THREADS]; // [level][thread] = (node, todo | depth << 8)
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const f3 ro = F3(p.pos[0], p.pos[1], p.pos[2]);
    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);

    int sp = 0;
    uint32_t node = 0, depth = 0; // depth of `node`
    f3 P = F3(0.f, 0.f, 0.f);     // corner of `node`
    float side = 0.5f;            // side of its children
    Accum<STRICT> acc;
    const uint32_t stack = stack_base(stack_mem);

    DfPlanes q;

#if XN_DF_NODE_AT_A_TIME
    if (!STRICT) {
        // Fast mode: the SET of leaves a ray adds (and every chord) is the shader's, but the order of
        // the additions is free (the mode's contract is <= 1/255, not bit-identity).  That removes the
        // return step altogether: on entering a node all of its hit leaf children are added at once
        // (child index known at compile time: no plane selection; the chord t_max - max(t_min, 0)
        // that decides the hit -- positive iff `t_min < t_max && t_max > 0` -- is the length added),
        // and only the hit INTERNAL children are remembered.  Going back to an ancestor then needs
        // its corner alone (to place the next internal child), not its planes -- so every iteration
        // of the loop enters exactly one node, with the whole warp on the plane arithmetic.
        for (;;) {
            df_planes(P, side, rrd, bias, q);
#if XN_LDG256
            uint4 w0, w1;
            ldg256(reinterpret_cast<const uint32_t*>(p.cnodes) + node, w0, w1);
#else
            const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(p.cnodes) + node));
            const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(p.cnodes) + node) + 1);
#endif
            const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            uint32_t inner = 0, hits = 0;
#pragma unroll
            for (uint32_t c = 0; c < 8u; ++c) {
                float t_min, t_max;
                df_child(q, c, t_min, t_max);
                const float chord = t_max - fmaxf(t_min, 0.0f);
                if (chord > 0.0f) {
                    if (STATS) ++hits;
                    if (word_is_leaf(w[c])) {
                        st.read(4); // color
                        acc.add(w[c], chord);
                    } else {
                        inner |= 1u << c;
                        if (c != 7u) st.read(4); // the shader's read on returning from this child (svo_df.comp:58)
                    }
                }
            }
            if (STATS) {
                st.steps += 8;
                st.read(32 + 4 * hits); // children[i] of every iteration + is_leaf_depth of the hit ones
            }
            // next node to enter: the first hit internal child, else the next one of the nearest
            // ancestor that has any left
            uint32_t c;
            if (inner != 0u) {
                c = (uint32_t)__ffs((int)inner) - 1u;
                inner &= inner - 1u;
                if (inner != 0u) {
                    stack_store(stack, (uint32_t)min(sp, LEVELS - 1), node, inner | (depth << 8));
                    ++sp;
                }
            } else {
                if (sp == 0) break;
                const uint2 e = stack_load(stack, (uint32_t)min(sp - 1, LEVELS - 1));
                node = e.x;
                uint32_t rest = e.y & 0xFFu;
                depth = e.y >> 8;
                c = (uint32_t)__ffs((int)rest) - 1u;
                rest &= rest - 1u;
                if (rest != 0u) stack_store(stack, (uint32_t)min(sp - 1, LEVELS - 1), node, rest | (depth << 8));
                else --sp;
                const float size = __int_as_float((127 - (int)depth) << 23); // exp2(-depth): the node's side
                side = size * 0.5f;
                const float rsize = pow2_reciprocal(size);
                P = F3(size * floorf(P.x * rsize), size * floorf(P.y * rsize), size * floorf(P.z * rsize));
            }
            if (c & 4u) P.x += side;
            if (c & 2u) P.y += side;
            if (c & 1u) P.z += side;
            node = load_word(p.cnodes, node, c);
            ++depth;
            side *= 0.5f;
        }
        store_result(p, ix, iy, acc.finish(voxel_emission_coeff(p, rd)), st);
        return;
    }
#endif
    df_planes(P, side, rrd, bias, q);
    uint32_t todo = df_hit_mask(q); // hit children of `node` not visited yet
    if (STATS) {
        st.steps += 8;
        st.read(32 + 4 * __popc(todo)); // children[i] of every iteration + is_leaf_depth of the hit ones
    }
#if XN_DF_TWO_PHASE
    for (;;) {
        // leaf run: hit children in index order (the shader's accumulation order) up to the first
        // internal one, whose word stays in `s`
        uint32_t s = 0, c = 0;
        bool internal = false;
        while (todo != 0u && !internal) {
            c = (uint32_t)__ffs((int)todo) - 1u;
            todo &= todo - 1u;
            s = load_word(p.cnodes, node, c);
            if (word_is_leaf(s)) {
                float t_min, t_max;
                df_child(q, c, t_min, t_max);
                st.read(4); // color
                acc.add(s, t_max - fmaxf(t_min, 0.0f));
            } else {
                internal = true;
            }
        }
        // node step
        if (internal) {
            // the shader pushes unless this is child 7 and pops when the subtree is done (one read of
            // nodes[node].is_leaf_depth per pop, svo_df.comp:58)
            if (c != 7u) st.read(4);
            const f3 Pc = F3(P.x + ((c & 4u) ? side : 0.0f), P.y + ((c & 2u) ? side : 0.0f), P.z + ((c & 1u) ? side : 0.0f));
            const float side_c = side * 0.5f;
#if XN_DF_INLINE
            // Most internal nodes a ray meets are parents of leaves only (7/8 of the internal nodes of
            // a tree sit on its last level).  Such a child is consumed IN PLACE: its planes go to a
            // second register set, its hit leaves are added, and the current node's state -- planes,
            // children left -- is untouched, so there is no stack entry, no return step and no
            // recomputation of the parent's planes.  Only when the child turns out to have an internal
            // child of its own does the traversal really move into it (what has been added so far
            // stays: the order of the shader's additions is kept).
            DfPlanes q2;
            df_planes(Pc, side_c, rrd, bias, q2);
            uint32_t todo2 = df_hit_mask(q2);
            if (STATS) {
                st.steps += 8;
                st.read(32 + 4 * __popc(todo2));
            }
            bool deep = false;
            while (todo2 != 0u && !deep) {
                const uint32_t c2 = (uint32_t)__ffs((int)todo2) - 1u;
                const uint32_t w = load_word(p.cnodes, s, c2);
                if (word_is_leaf(w)) {
                    todo2 &= todo2 - 1u;
                    float t_min, t_max;
                    df_child(q2, c2, t_min, t_max);
                    st.read(4); // color
                    acc.add(w, t_max - fmaxf(t_min, 0.0f));
                } else {
                    deep = true; // c2 stays in todo2: the leaf run of the entered node meets it again
                }
            }
            if (!deep) continue;
            if (todo != 0u) {
                stack_store(stack, (uint32_t)min(sp, LEVELS - 1), node, todo | (depth << 8));
                ++sp;
            }
            P = Pc;
            node = s;
            ++depth;
            side = side_c;
            q = q2;
            todo = todo2;
            continue;
#else
            if (todo != 0u) {
                stack_store(stack, (uint32_t)min(sp, LEVELS - 1), node, todo | (depth << 8));
                ++sp;
            }
            P = Pc;
            node = s;
            ++depth;
            side = side_c;
#endif
        } else {
            // node finished: back to the nearest ancestor with children left
            if (sp == 0) break;
            --sp;
            const uint2 e = stack_load(stack, (uint32_t)min(sp, LEVELS - 1));
            node = e.x;
            todo = e.y & 0xFFu;
            depth = e.y >> 8;
            const float size = __int_as_float((127 - (int)depth) << 23); // exp2(-depth): the node's side
            side = size * 0.5f;
            // corner of the node: the corner of any descendant rounded down to a multiple of its size
            // (levels left by tail descents pushed nothing, so there is no per-level offset to undo)
            const float rsize = pow2_reciprocal(size);
            P = F3(size * floorf(P.x * rsize), size * floorf(P.y * rsize), size * floorf(P.z * rsize));
        }
        df_planes(P, side, rrd, bias, q);
        if (internal) {
            todo = df_hit_mask(q);
            if (STATS) {
                st.steps += 8;
                st.read(32 + 4 * __popc(todo));
            }
        }
    }
#else
    for (;;) {
        if (todo == 0u) {
            // node finished: back to the nearest ancestor with children left
            if (sp == 0) break;
            --sp;
            const uint2 e = stack_load(stack, (uint32_t)min(sp, LEVELS - 1));
            node = e.x;
            todo = e.y & 0xFFu;
            depth = e.y >> 8;
            const float size = __int_as_float((127 - (int)depth) << 23); // exp2(-depth): the node's side
            side = size * 0.5f;
            const float rsize = pow2_reciprocal(size);
            P = F3(size * floorf(P.x * rsize), size * floorf(P.y * rsize), size * floorf(P.z * rsize));
            df_planes(P, side, rrd, bias, q);
            continue;
        }
        const uint32_t c = (uint32_t)__ffs((int)todo) - 1u;
        todo &= todo - 1u;
        const uint32_t s = load_word(p.cnodes, node, c);
        if (word_is_leaf(s)) {
            float t_min, t_max;
            df_child(q, c, t_min, t_max);
            st.read(4); // color
            acc.add(s, t_max - fmaxf(t_min, 0.0f));
        } else {
            if (c != 7u) st.read(4);
            if (todo != 0u) {
                stack_store(stack, (uint32_t)min(sp, LEVELS - 1), node, todo | (depth << 8));
                ++sp;
            }
            if (c & 4u) P.x += side;
            if (c & 2u) P.y += side;
            if (c & 1u) P.z += side;
            node = s;
            ++depth;
            side *= 0.5f;
            df_planes(P, side, rrd, bias, q);
            todo = df_hit_mask(q);
            if (STATS) {
                st.steps += 8;
                st.read(32 + 4 * __popc(todo));
            }
        }
    }
#endif
    store_result(p, ix, iy, acc.finish(voxel_emission_coeff(p, rd)), st);
}

// ---------------------------------------------------------------------------------
// Persistent-thread ray pool.  p.pool[0] counts the rays handed out in this launch (zeroed by the
// launcher); ray id -> pixel keeps the static launch's shape: 32 consecutive ids are one warp's
// 8x4 tile, BLOCK_THREADS ids one 16-row block, blocks row-major over the owned stripes.
// Idle lanes of a warp are served together: one atomic for the warp, offsets from the ballot.
// Returns true on the lanes that received a ray inside the frame; `dry` turns true (for the
// whole warp) once the counter has passed the last ray.
// ---------------------------------------------------------------------------------
#ifndef XN_POOL_MIN_ACTIVE
#define XN_POOL_MIN_ACTIVE 24
#endif
__device__ __forceinline__ bool ray_pool_draw(const FrameParams& p, bool idle, bool& dry, uint32_t& ix, uint32_t& iy) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, idle);
    if (m == 0u || dry) return false;
    const uint32_t lane = threadIdx.x & 31u, n = (uint32_t)__popc(m);
    const int leader = __ffs((int)m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(p.pool, n);
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    dry = base + n >= p.pool_total;
    if (!idle) return false;
    const uint32_t id = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    if (id >= p.pool_total) return false;
    const uint32_t blk = id / (uint32_t)BLOCK_THREADS, in = id % (uint32_t)BLOCK_THREADS, w = in >> 5, l = in & 31u;
    const uint32_t bx = blk % p.pool_blocks_x, by = blk / p.pool_blocks_x;
    ix = bx * BLOCK_W + (w % BLOCK_WARPS_X) * TILE_W + l % TILE_W;
    iy = (by * p.il_count + p.il_index) * BLOCK_H + (w / BLOCK_WARPS_X) * TILE_H + l / TILE_W;
    return ix < p.out_w && iy < p.out_h;
}

// ESVO child selection.  The compiler turns `if (c) { pos += d; idx ^= bit; }` into compare +
// FADD + FSEL + SEL chains on the ALU pipe, which bounded this kernel (ncu: ALU 70 %, issue 72 %).
// XN_ESVO_PTX = 2 writes the comparisons as 1.0 / 0.0 flags (`set`) followed by exact FMAs, which
// run on the FMA pipe; 0 is the plain C++.  (Predicated `@p add` in PTX was tried too: ptxas turns
// it back into FADD + FSEL.)
#ifndef XN_ESVO_PTX
#define XN_ESVO_PTX 2
#endif
// ADVANCE (esvo.comp:96-104): axes whose corner time equals tc_max step back by `se`;
// returns the step mask (x = 4, y = 2, z = 1)
__device__ __forceinline__ uint32_t esvo_advance(float tcorx, float tcory, float tcorz, float tc_max, float se,
                                                 float& posx, float& posy, float& posz, float& fx, float& fy,
                                                 float& fz) {
#if XN_ESVO_PTX == 2
    // comparison results as 1.0f / 0.0f (one FSET each); the position update and the step mask are
    // then exact FMAs on the FMA pipe (pos - 1.0 * se, 4 ax + 2 ay + az) and one float -> int
    // conversion, instead of predicate -> SEL / FSEL chains on the ALU pipe
    uint32_t m;
    asm("{\n\t"
        ".reg .f32 mf;\n\t"
        "set.le.f32.f32 %4, %7, %10;\n\t"
        "set.le.f32.f32 %5, %8, %10;\n\t"
        "set.le.f32.f32 %6, %9, %10;\n\t"
        "fma.rn.f32 %1, %4, %11, %1;\n\t"
        "fma.rn.f32 %2, %5, %11, %2;\n\t"
        "fma.rn.f32 %3, %6, %11, %3;\n\t"
        "fma.rn.f32 mf, %4, 0f40800000, %6;\n\t"
        "fma.rn.f32 mf, %5, 0f40000000, mf;\n\t"
        "cvt.rzi.u32.f32 %0, mf;\n\t"
        "}"
        : "=r"(m), "+f"(posx), "+f"(posy), "+f"(posz), "=&f"(fx), "=&f"(fy), "=&f"(fz)
        : "f"(tcorx), "f"(tcory), "f"(tcorz), "f"(tc_max), "f"(-se));
    return m;
#else
    const bool ax = tcorx <= tc_max, ay = tcory <= tc_max, az = tcorz <= tc_max;
    if (ax) posx -= se;
    if (ay) posy -= se;
    if (az) posz -= se;
    fx = ax ? 1.0f : 0.0f, fy = ay ? 1.0f : 0.0f, fz = az ? 1.0f : 0.0f;
    return (ax ? 4u : 0u) | (ay ? 2u : 0u) | (az ? 1u : 0u);
#endif
}
// PUSH child selection (esvo.comp:84-90): the half of the node the ray enters first on every axis
__device__ __forceinline__ uint32_t esvo_first_child(float tcenx, float tceny, float tcenz, float t_min, float se,
                                                     float& posx, float& posy, float& posz) {
#if XN_ESVO_PTX == 2
    uint32_t m;
    asm("{\n\t"
        ".reg .f32 ax, ay, az, mf;\n\t"
        "set.gt.f32.f32 ax, %4, %7;\n\t"
        "set.gt.f32.f32 ay, %5, %7;\n\t"
        "set.gt.f32.f32 az, %6, %7;\n\t"
        "fma.rn.f32 %1, ax, %8, %1;\n\t"
        "fma.rn.f32 %2, ay, %8, %2;\n\t"
        "fma.rn.f32 %3, az, %8, %3;\n\t"
        "fma.rn.f32 mf, ax, 0f40800000, az;\n\t"
        "fma.rn.f32 mf, ay, 0f40000000, mf;\n\t"
        "cvt.rzi.u32.f32 %0, mf;\n\t"
        "}"
        : "=r"(m), "+f"(posx), "+f"(posy), "+f"(posz)
        : "f"(tcenx), "f"(tceny), "f"(tcenz), "f"(t_min), "f"(se));
    return m;
#else
    uint32_t m = 0;
    if (tcenx > t_min) { m ^= 4u; posx += se; }
    if (tceny > t_min) { m ^= 2u; posy += se; }
    if (tcenz > t_min) { m ^= 1u; posz += se; }
    return m;
#endif
}

// Leaf brick (fast mode only).  The child `s` of the current node is an internal node whose eight
// children are all leaves: instead of PUSHing into it and visiting its children one loop iteration
// each (two to four iterations of PUSH / ADVANCE / POP at the lane divergence of the deepest level,
// where neighbouring rays are in different phases all the time), its share of the integral is taken
// here in closed form.  In the mirrored frame the ray leaves the upper half of every axis when it
// crosses the node's centre plane at tcen_i = (pos_i + half) tc_i - tb_i; clipped to the ray's stay
// [t_min, t_end] in the node and sorted, the three crossings cut the stay into at most four chords,
// and the child that holds a chord has its bit set on exactly the axes whose crossing is still
// ahead.  These are the chords esvo.comp:65-104 walks through (a child's exit time is the next
// crossing), up to the rounding of t -- which is why the strict mode keeps the loop.  Chords of
// length zero (a plane crossed outside the stay, or two at once) add colour x 0.
// Measured (profiles/README.md, round 2 session 4): on the bunny tree, script frame 120, 46 % fewer loop
// iterations, 19 % fewer warp instructions, 21.2 -> 23.1 lanes per instruction, kernel time -13 %;
// whole camera path +4 % (1203 -> 1252 Mrays/s).  On the 2048^3 TNG tree -20 %: there few lanes of a
// warp stand before a brick at the same time, and the 80 instructions of the closed form are paid
// per warp while the iterations they replace were shared with the other lanes' iterations (taking
// the closed form only when a ballot finds enough such lanes costs more than it saves: 939).  So
// this is an opt-in (XN_ESVO_BRICKS=1), like the ray pool.
#ifndef XN_ESVO_BRICK_MIN_BLOCKS
#define XN_ESVO_BRICK_MIN_BLOCKS 4
#endif
template <bool STRICT>
__device__ __forceinline__ void esvo_leaf_brick(const CNode* __restrict__ nodes, uint32_t s, uint32_t octant_mask,
                                                float tcx, float tcy, float tcz, float tcorx, float tcory, float tcorz,
                                                float half, float t_min, float t_end, Accum<STRICT>& acc) {
    const float ax = fminf(fmaxf(__fmaf_rn(half, tcx, tcorx), t_min), t_end);
    const float ay = fminf(fmaxf(__fmaf_rn(half, tcy, tcory), t_min), t_end);
    const float az = fminf(fmaxf(__fmaf_rn(half, tcz, tcorz), t_min), t_end);
    const float lo = fminf(ax, ay), hi = fmaxf(ax, ay);
    const float t1 = fminf(lo, az), t3 = fmaxf(hi, az), t2 = fmaxf(lo, fminf(hi, az));
    // child holding the chord that starts at ts: upper half on the axes not yet crossed
#define XN_BRICK_CHILD(ts) ((ax > (ts) ? 4u : 0u) | (ay > (ts) ? 2u : 0u) | (az > (ts) ? 1u : 0u))
    const uint32_t w0 = load_word(nodes, s, XN_BRICK_CHILD(t_min) ^ octant_mask);
    const uint32_t w1 = load_word(nodes, s, XN_BRICK_CHILD(t1) ^ octant_mask);
    const uint32_t w2 = load_word(nodes, s, XN_BRICK_CHILD(t2) ^ octant_mask);
    const uint32_t w3 = load_word(nodes, s, octant_mask); // every plane crossed: the lowest child
#undef XN_BRICK_CHILD
    acc.add(w0, t1 - t_min);
    acc.add(w1, t2 - t1);
    acc.add(w2, t3 - t2);
    acc.add(w3, t_end - t3);
}

// ---------------------------------------------------------------------------------
// esvo (resources/esvo.comp:10-157)
// The reference indexes its stacks by `scale` (22 downwards); here level = 22 - scale, so
// LEVELS (>= tree depth) levels of shared memory are enough.
// ---------------------------------------------------------------------------------
//
// POOL (north_star's persistent-thread ray pool, after Aila & Laine 2009): the grid is one wave of
// resident blocks whose warps draw rays from a per-launch counter (ray_pool_draw: one atomic per
// warp and refill, lane offsets from the ballot of idle lanes) until it runs dry.  A warp leaves
// the traversal loop to refill when fewer than XN_POOL_MIN_ACTIVE of its lanes still have a ray;
// lanes that kept theirs re-enter the loop with their state untouched (the stack is per thread, not
// per ray slot).  Rays are numbered so that 32 consecutive ones are an 8x4 tile and 256 a block of
// the static launch: a freshly filled warp is as coherent as a static one.  What it buys is
// measured, not assumed: profiles/README.md (ray pool).
// BRICKS: child words at or above p.brick_base are leaf bricks (esvo_leaf_brick); a separate
// instantiation, so trees without them run the plain loop's code unchanged.
template <bool STATS, bool STRICT, int LEVELS, bool POOL = false, bool BRICKS = false>
__global__ void __launch_bounds__(BLOCK_THREADS, BRICKS ? XN_ESVO_BRICK_MIN_BLOCKS : XN_ESVO_MIN_BLOCKS)
    esvo_kernel(const __grid_constant__ FrameParams p) {
    // [scale][thread] = (parent, bits(t_max)).  Trees this instantiation is launched for are shallower
    // than LEVELS; the index is clamped instead of bounds-tested, so a malformed file (a cycle of
    // child pointers) aliases its own thread's last entry and nothing else.  Shared memory is
    // kept small on purpose: what the stacks take is carved out of the L1 cache.
    __shared__ uint2 stack_mem[LEVELS * BLOCK_THREADS];
    const uint32_t cast_stack_depth = 23u;
    uint32_t ix = 0, iy = 0;
    if (!POOL) {
        thread_pixel(p, ix, iy);
        if (ix >= p.out_w || iy >= p.out_h) return;
    }
    RayStats<STATS> st;
    // POOL: everything a ray carries is declared out here so that a lane keeps it across refills
    f3 rd = F3(0.f, 0.f, 0.f);
    float tcx = 0.f, tcy = 0.f, tcz = 0.f, tbx = 0.f, tby = 0.f, tbz = 0.f, t_min = 0.f, t_max = 0.f;
    float posx = 1.f, posy = 1.f, posz = 1.f, scale_exp2 = 0.5f;
    uint32_t octant_mask = 0, parent = 0, idx = 0, s = 0;
    uint32_t scale = cast_stack_depth; // >= cast_stack_depth: this lane has no ray
#if !XN_ESVO_ALWAYS_STORE
    float h = 0.f;
#endif
    Accum<STRICT> acc;
    bool dry = false;  // POOL: the counter has passed the last ray (warp-uniform)
    bool have = false; // POOL: this lane holds a ray
    constexpr uint32_t MIN_SCALE = LEVELS >= 23 ? 0u : 23u - (uint32_t)LEVELS;
    const uint32_t stack = stack_base(stack_mem) - MIN_SCALE * (uint32_t)(BLOCK_THREADS * sizeof(uint2));
    const CNode* __restrict__ nodes = p.cnodes;
  for (;;) {
    bool fresh = !POOL;
    if (POOL) {
        fresh = ray_pool_draw(p, !have, dry, ix, iy);
        if (__all_sync(0xFFFFFFFFu, !have && !fresh)) {
            if (dry) return;
            continue; // every ray drawn fell outside the frame: draw again
        }
    }
    if (fresh) {
    if (POOL) {
        st = RayStats<STATS>();
        acc = Accum<STRICT>();
        have = true;
    }
    rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    f3 ro = F3(p.pos[0] + 1.0f, p.pos[1] + 1.0f, p.pos[2] + 1.0f);
    {
        // aabb_intersect(vec3(1), vec3(2), ro, rd), esvo.comp:10-21
        const f3 r = F3(1.0f / (rd.x + 0.00000001f), 1.0f / (rd.y + 0.00000001f), 1.0f / (rd.z + 0.00000001f));
        const f3 tbot = F3((1.0f - ro.x) * r.x, (1.0f - ro.y) * r.y, (1.0f - ro.z) * r.z);
        const f3 ttop = F3((2.0f - ro.x) * r.x, (2.0f - ro.y) * r.y, (2.0f - ro.z) * r.z);
        const f3 tmn = F3(gmin(ttop.x, tbot.x), gmin(ttop.y, tbot.y), gmin(ttop.z, tbot.z));
        const float t0 = gmax(gmax(tmn.x, tmn.y), gmax(tmn.x, tmn.z));
        const float adv = gmax(t0, 0.0f);
        ro = F3(ro.x + adv * rd.x, ro.y + adv * rd.y, ro.z + adv * rd.z);
    }

    tcx = 1.0f / -fabsf(rd.x), tcy = 1.0f / -fabsf(rd.y), tcz = 1.0f / -fabsf(rd.z);
    tbx = tcx * ro.x, tby = tcy * ro.y, tbz = tcz * ro.z;
    octant_mask = 0;
    if (rd.x > 0.0f) { tbx = 3.0f * tcx - tbx; octant_mask ^= 4u; }
    if (rd.y > 0.0f) { tby = 3.0f * tcy - tby; octant_mask ^= 2u; }
    if (rd.z > 0.0f) { tbz = 3.0f * tcz - tbz; octant_mask ^= 1u; }

    t_min = max_elem(F3(2.0f * tcx - tbx, 2.0f * tcy - tby, 2.0f * tcz - tbz));
    t_max = min_elem(F3(tcx - tbx, tcy - tby, tcz - tbz));
#if !XN_ESVO_ALWAYS_STORE
    h = t_max;
#endif
    t_min = gmax(t_min, 0.0f);
    t_max = gmin(t_max, sqrtf(3.0f));

    parent = 0, idx = 0;
    posx = 1.f, posy = 1.f, posz = 1.f;
    scale = cast_stack_depth - 1u;
    scale_exp2 = 0.5f;
    if (1.5f * tcx - tbx > t_min) { posx = 1.5f; idx ^= 4u; }
    if (1.5f * tcy - tby > t_min) { posy = 1.5f; idx ^= 2u; }
    if (1.5f * tcz - tbz > t_min) { posz = 1.5f; idx ^= 1u; }

    // entries are indexed by `scale` itself, as in esvo.comp:34-35: the base is moved down by the
    // (23 - LEVELS) scales this instantiation never reaches, so a push / pop address is one scaled add

    // The child descriptor of (parent, idx) is requested at the END of the previous iteration,
    // at a single load site (99.6 % of iterations consume it, ncu r01), so the t_corner
    // arithmetic of the next iteration overlaps the load instead of waiting behind it.
    s = load_word(nodes, parent, idx ^ octant_mask);
    } // ray set-up

    if (!POOL || have) {
    while (scale < cast_stack_depth) {
        // POOL: too few lanes left with a ray -> out to the refill (this lane keeps its state)
        if (POOL && !dry && __popc(__activemask()) < XN_POOL_MIN_ACTIVE) break;
        st.step();
        const float tcorx = posx * tcx - tbx, tcory = posy * tcy - tby, tcorz = posz * tcz - tbz;
        const float tc_max = fminf(tcorx, fminf(tcory, tcorz));

        bool pushed = false;
        {
            // esvo.comp:70-72 tests t_min <= t_max, then t_min <= tv_max = min(t_max, tc_max); the
            // second test implies the first (no NaNs on this path), so one comparison decides
            const float tv_max = fminf(t_max, tc_max);
            if (t_min <= tv_max) {
                st.read(8); // children[idx ^ octant_mask] + nodes[child].is_leaf_depth
                if (word_is_leaf(s)) {
                    st.read(4); // color
                    acc.add(s, tv_max - t_min);
                } else if (BRICKS && s >= p.brick_base) {
                    // eight leaves below: integrated here, then ADVANCE as after a leaf
                    esvo_leaf_brick(nodes, s, octant_mask, tcx, tcy, tcz, tcorx, tcory, tcorz, scale_exp2 * 0.5f, t_min,
                                    tv_max, acc);
                } else {
                    // PUSH.  esvo.comp:79-82 skips the stack write when the child's exit time is not
                    // below `h` (the parent will never be popped to); writing always stores the
                    // same (parent, t_max) the reference would have stored at this level, is never
                    // observable, and saves the comparison and the bookkeeping of h.
#if XN_ESVO_ALWAYS_STORE
                    stack_store(stack, max(scale, MIN_SCALE), parent,
                                __float_as_uint(t_max));
#else
                    if (tc_max < h)
                        stack_store(stack, max(scale, MIN_SCALE), parent,
                                    __float_as_uint(t_max));
                    h = tc_max;
#endif
                    parent = s;
                    --scale;
                    scale_exp2 *= 0.5f;
                    const float tcenx = scale_exp2 * tcx + tcorx, tceny = scale_exp2 * tcy + tcory,
                                tcenz = scale_exp2 * tcz + tcorz;
                    idx = esvo_first_child(tcenx, tceny, tcenz, t_min, scale_exp2, posx, posy, posz);
                    t_max = tv_max;
                    pushed = true;
                }
            }
        }

        if (!pushed) {
            // ADVANCE
            float fx, fy, fz; // 1.0 on the axes that stepped
            const uint32_t step_mask = esvo_advance(tcorx, tcory, tcorz, tc_max, scale_exp2, posx, posy, posz, fx, fy, fz);
            const bool ax = (step_mask & 4u) != 0u, ay = (step_mask & 2u) != 0u, az = (step_mask & 1u) != 0u;
            t_min = tc_max;
            idx ^= step_mask;

            if ((idx & step_mask) != 0u) {
                // POP
#if XN_ESVO_POP
                // pos + scale_exp2 is the position before the step, exactly; axes that did not
                // step contribute 0 (esvo.comp:108-116)
                const float ox = __fmaf_rn(fx, scale_exp2, posx), oy = __fmaf_rn(fy, scale_exp2, posy),
                            oz = __fmaf_rn(fz, scale_exp2, posz);
                const uint32_t dbits = (__float_as_uint(ox) ^ __float_as_uint(posx)) |
                                       (__float_as_uint(oy) ^ __float_as_uint(posy)) |
                                       (__float_as_uint(oz) ^ __float_as_uint(posz));
#else
                uint32_t dbits = 0;
                if (ax) dbits |= __float_as_uint(posx) ^ __float_as_uint(posx + scale_exp2);
                if (ay) dbits |= __float_as_uint(posy) ^ __float_as_uint(posy + scale_exp2);
                if (az) dbits |= __float_as_uint(posz) ^ __float_as_uint(posz + scale_exp2);
#endif
                // esvo.comp:117 takes the exponent of float(dbits); dbits is a union of carry runs
                // of at most 23 bits inside the cube (exact in binary32), so the index of its
                // highest set bit is the same number, and anything >= 23 (or dbits == 0) leaves
                scale = 31u - (uint32_t)__clz((int)dbits);
                if (scale >= cast_stack_depth) break; // left the cube (also guards the reference's
                                                      // underflowed stack read, esvo.comp:119-123)
                scale_exp2 = __uint_as_float(scale * 0x00800000u + 0x34000000u); // exp2(scale - 23): one IMAD
                const uint2 e = stack_load(stack, max(scale, MIN_SCALE));
#if !XN_ESVO_ALWAYS_STORE
                h = 0.0f;
#endif
                parent = e.x;
                t_max = __uint_as_float(e.y);
                const uint32_t shx = __float_as_uint(posx) >> scale, shy = __float_as_uint(posy) >> scale,
                               shz = __float_as_uint(posz) >> scale;
#if XN_ESVO_POP
                const uint32_t keep = 0xFFFFFFFFu << scale;
                posx = __uint_as_float(__float_as_uint(posx) & keep);
                posy = __uint_as_float(__float_as_uint(posy) & keep);
                posz = __uint_as_float(__float_as_uint(posz) & keep);
#else
                posx = __uint_as_float(shx << scale);
                posy = __uint_as_float(shy << scale);
                posz = __uint_as_float(shz << scale);
#endif
                idx = (shx & 1u) * 4u + (shy & 1u) * 2u + (shz & 1u);
            }
        }
        s = load_word(nodes, parent, idx ^ octant_mask);
    }
    if (!POOL || scale >= cast_stack_depth) {
        store_result(p, ix, iy, acc.finish(voxel_emission_coeff(p, rd)), st);
        have = false;
    }
    }
    if (!POOL) return;
  }
}

// ---------------------------------------------------------------------------------
// svo_rope (resources/svo_rope.comp:50-153)
// ---------------------------------------------------------------------------------
template <bool STATS, bool STRICT>
__global__ void __launch_bounds__(BLOCK_THREADS, XN_SVO_MIN_BLOCKS) svo_rope_kernel(const __grid_constant__ FrameParams p) {
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const f3 ro = F3(p.pos[0], p.pos[1], p.pos[2]);

    f3 sgn = F3(gsign(rd.x), gsign(rd.y), gsign(rd.z));
    const uint32_t nbx = 1u - (uint32_t)gmax(sgn.x, 0.0f);
    const uint32_t nby = 3u - (uint32_t)gmax(sgn.y, 0.0f);
    const uint32_t nbz = 5u - (uint32_t)gmax(sgn.z, 0.0f);
    sgn = F3(sgn.x + 0.1f, sgn.y + 0.1f, sgn.z + 0.1f);

    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);

    Accum<STRICT> acc;
    float t_min, t_max;
    if (unit_cube_slab(rrd, bias, t_min, t_max)) {
        f3 pos = F3(ro.x + t_min * rd.x, ro.y + t_min * rd.y, ro.z + t_min * rd.z);
        uint32_t node = 0, meta = p.root_meta;
        f3 offset = F3(0.f, 0.f, 0.f);
        float side = 1.0f;
        descend(p.nodes, pos, node, meta, offset, side, st);

        for (;;) {
            float u_min, u_max;
            f3 far;
            node_slab(offset, side, rrd, bias, u_min, u_max, far);
            const float step = u_max - (XN_SLAB_FMNMX ? fmaxf(u_min, 0.0f) : gmax(u_min, 0.0f));
            st.read(4); // color
            acc.add(meta, step);
            st.step();

            // neighbor_index, svo_rope.comp:50-63 (ties go to the later axis)
            uint32_t n;
            if (far.x < fminf(far.y, far.z)) {
                n = nbx;
                offset.x += sgn.x * side;
            } else if (far.y < far.z) {
                n = nby;
                offset.y += sgn.y * side;
            } else {
                n = nbz;
                offset.z += sgn.z * side;
            }
            st.read(4); // rope
            const uint2 s = load_slot(p.nodes, node, n);
            node = s.x;
            meta = s.y;
            if (node == 0u) break;

            // find_relative, svo_rope.comp:29-48
            pos = F3(ro.x + u_max * rd.x, ro.y + u_max * rd.y, ro.z + u_max * rd.z);
            side = __int_as_float((127 - (int)meta_depth(meta)) << 23); // exp2(-depth)
            const float rside = pow2_reciprocal(side);
            // offset -= mod(offset, side) (svo_rope.comp:38).  k = floor(offset / side) is an integer,
            // side * k is exact, and so are offset - side * k (a multiple of ulp(offset) below offset)
            // and offset minus that: the result is side * k itself, three operations instead of five
            offset.x = side * floorf(offset.x * rside);
            offset.y = side * floorf(offset.y * rside);
            offset.z = side * floorf(offset.z * rside);
            descend(p.nodes, pos, node, meta, offset, side, st);
        }
    }
    store_result(p, ix, iy, acc.finish(voxel_emission_coeff(p, rd)), st);
}

// ---------------------------------------------------------------------------------
// svo_rope over the 32-byte rope records (RNode, xn_device.cuh): the same walk, same arithmetic,
// same per-ray read counts as svo_rope_kernel; what differs is where a node's words live.  A visit
// to a leaf fetches its ONE sector as two 128-bit loads: colour (w[6]) and all six ropes, of which
// the exit face picks one.  The record a rope leads to is fetched whole as well; its tag word says
// whether it is the next leaf or an internal node to descend through (child words carry the
// child's leaf flag, so every further level is again one record fetch).
// ---------------------------------------------------------------------------------
struct RRec {
    uint4 a, b; // w[0..3], w[4..7]
};
__device__ __forceinline__ RRec load_rrec(const RNode* __restrict__ nodes, uint32_t i) {
    RRec r;
#if XN_LDG256
    ldg256(nodes + i, r.a, r.b);
#else
    const uint4* q = reinterpret_cast<const uint4*>(nodes + i);
    r.a = __ldg(q);
    r.b = __ldg(q + 1);
#endif
    return r;
}
__device__ __forceinline__ uint32_t rrec_word(const RRec& r, uint32_t k) { // k in 0..7, runtime
    const uint32_t lo = (k & 2u) ? ((k & 1u) ? r.a.w : r.a.z) : ((k & 1u) ? r.a.y : r.a.x);
    const uint32_t hi = (k & 2u) ? ((k & 1u) ? r.b.w : r.b.z) : ((k & 1u) ? r.b.y : r.b.x);
    return (k & 4u) ? hi : lo;
}
// find() / find_relative() loop body (svo_rope.comp:14-26, :33-47) from record `rec` of node `node`
// at `offset` / `extent`, down to the leaf containing pos; returns with the leaf's record in `rec`
template <bool STATS>
__device__ __forceinline__ void descend_rrec(const RNode* __restrict__ nodes, f3 pos, uint32_t& node, RRec& rec,
                                             f3& offset, float& extent, RayStats<STATS>& st) {
    bool leaf = rec.b.w == RNODE_LEAF_TAG;
    for (;;) {
        st.read(4); // is_leaf_depth
        if (leaf) return;
        extent *= 0.5f;
        uint32_t child;
        {
            const float cx = offset.x + extent, cy = offset.y + extent, cz = offset.z + extent;
            asm("{\n\t"
                ".reg .f32 fx, fy, fz, mf;\n\t"
                "set.ge.f32.f32 fx, %4, %7;\n\t"
                "set.ge.f32.f32 fy, %5, %8;\n\t"
                "set.ge.f32.f32 fz, %6, %9;\n\t"
                "fma.rn.f32 %1, fx, %10, %1;\n\t"
                "fma.rn.f32 %2, fy, %10, %2;\n\t"
                "fma.rn.f32 %3, fz, %10, %3;\n\t"
                "fma.rn.f32 mf, fx, 0f40800000, fz;\n\t"
                "fma.rn.f32 mf, fy, 0f40000000, mf;\n\t"
                "cvt.rzi.u32.f32 %0, mf;\n\t"
                "}"
                : "=r"(child), "+f"(offset.x), "+f"(offset.y), "+f"(offset.z)
                : "f"(pos.x), "f"(pos.y), "f"(pos.z), "f"(cx), "f"(cy), "f"(cz), "f"(extent));
        }
        st.read(4); // children[child]
        const uint32_t w = rrec_word(rec, child);
        node = w & 0x7FFFFFFFu;
        leaf = (w >> 31) != 0u;
        rec = load_rrec(nodes, node);
    }
}

template <bool STATS, bool STRICT>
__global__ void __launch_bounds__(BLOCK_THREADS, XN_SVO_MIN_BLOCKS) svo_rope32_kernel(const __grid_constant__ FrameParams p) {
    uint32_t ix, iy;
    thread_pixel(p, ix, iy);
    if (ix >= p.out_w || iy >= p.out_h) return;
    RayStats<STATS> st;

    const f3 rd = make_ray(p, p.out_x + (int32_t)ix, p.out_y + (int32_t)iy);
    const f3 ro = F3(p.pos[0], p.pos[1], p.pos[2]);

    f3 sgn = F3(gsign(rd.x), gsign(rd.y), gsign(rd.z));
    const uint32_t nbx = 1u - (uint32_t)gmax(sgn.x, 0.0f);
    const uint32_t nby = 3u - (uint32_t)gmax(sgn.y, 0.0f);
    const uint32_t nbz = 5u - (uint32_t)gmax(sgn.z, 0.0f);
    sgn = F3(sgn.x + 0.1f, sgn.y + 0.1f, sgn.z + 0.1f);

    const f3 rrd = F3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
    const f3 bias = F3(rrd.x * ro.x, rrd.y * ro.y, rrd.z * ro.z);
    const RNode* __restrict__ nodes = p.rnodes;

    Accum<STRICT> acc;
    float t_min, t_max;
    if (unit_cube_slab(rrd, bias, t_min, t_max)) {
        f3 pos = F3(ro.x + t_min * rd.x, ro.y + t_min * rd.y, ro.z + t_min * rd.z);
        uint32_t node = 0;
        f3 offset = F3(0.f, 0.f, 0.f);
        float side = 1.0f;
        RRec rec = load_rrec(nodes, 0u);
        descend_rrec(nodes, pos, node, rec, offset, side, st);

        for (;;) {
            float u_min, u_max;
            f3 far;
            node_slab(offset, side, rrd, bias, u_min, u_max, far);
            const float step = u_max - (XN_SLAB_FMNMX ? fmaxf(u_min, 0.0f) : gmax(u_min, 0.0f));
            st.read(4); // color
            acc.add(rec.b.z, step);
            st.step();

            // neighbor_index, svo_rope.comp:50-63 (ties go to the later axis): the three candidate
            // ropes of this ray are fixed by its signs
            uint32_t r;
            if (far.x < fminf(far.y, far.z)) {
                r = nbx ? rec.a.y : rec.a.x;
                offset.x += sgn.x * side;
            } else if (far.y < far.z) {
                r = nby == 3u ? rec.a.w : rec.a.z;
                offset.y += sgn.y * side;
            } else {
                r = nbz == 5u ? rec.b.y : rec.b.x;
                offset.z += sgn.z * side;
            }
            st.read(4); // rope
            if (r == 0u) break;
            node = r & 0x0FFFFFFFu;
            rec = load_rrec(nodes, node);

            // find_relative, svo_rope.comp:29-48
            pos = F3(ro.x + u_max * rd.x, ro.y + u_max * rd.y, ro.z + u_max * rd.z);
            side = __int_as_float((127 - (int)(r >> 28)) << 23); // exp2(-depth of the neighbour)
            const float rside = pow2_reciprocal(side);
            // offset -= mod(offset, side) as side * floor(offset / side) (exact, see svo_rope_kernel)
            offset.x = side * floorf(offset.x * rside);
            offset.y = side * floorf(offset.y * rside);
            offset.z = side * floorf(offset.z * rside);
            descend_rrec(nodes, pos, node, rec, offset, side, st);
        }
    }
    store_result(p, ix, iy, acc.finish(voxel_emission_coeff(p, rd)), st);
}

// ---------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------
static bool force_idx64() {
    static const bool v = [] {
        const char* e = getenv("XN_FORCE_IDX64");
        return e && e[0] == '1';
    }();
    return v;
}

// XN_RAY_POOL=1: the ESVO draws its rays from the persistent pool (measured: profiles/README.md)
static bool ray_pool_mode() {
    const char* e = getenv("XN_RAY_POOL"); // read per launch: tests and A/B runs flip it inside one process
    return e && e[0] == '1';
}

template <bool STATS, bool STRICT>
static cudaError_t launch_t(int traversal, const FrameParams& p, cudaStream_t stream) {
    const uint32_t stripes = (p.out_h + BLOCK_H - 1) / BLOCK_H;
    if (p.il_count == 0 || p.il_index >= p.il_count) return cudaErrorInvalidValue;
    const uint32_t owned = stripes > p.il_index ? (stripes - p.il_index + p.il_count - 1) / p.il_count : 0;
    if (owned == 0) return cudaSuccess;
    const dim3 grid((p.out_w + BLOCK_W - 1) / BLOCK_W, owned, 1);
    const dim3 block(BLOCK_THREADS, 1, 1);
    const bool deep = p.max_depth + 1u > 12u; // 12 stack levels cover trees up to 2048^3
    switch (traversal) {
        case 0:
            // XN_FORCE_IDX64=1 (test knob) runs the 64-bit-index kernel on small grids too
            if (p.tex_unorm != 0ull) { // texture residency
                if (p.skip_table) dda_skip_tex_kernel<STATS, STRICT><<<grid, block, 0, stream>>>(p);
                else dda_tex_kernel<STATS, STRICT><<<grid, block, 0, stream>>>(p);
            } else if (p.bk_slots != 0) { // bricked residency (xn_brick.h)
                if (p.bk_slots <= (1ull << 32) && !force_idx64())
                    dda_kernel<STATS, STRICT, BrickCursor<false>><<<grid, block, 0, stream>>>(p);
                else if (p.bk_top == 2u && p.bk_hs[2] <= 32u)
                    dda_kernel<STATS, STRICT, BrickCursor<true>><<<grid, block, 0, stream>>>(p);
                else
                    return cudaErrorInvalidValue;
            } else if ((uint64_t)p.nx * p.ny * p.nz < (1ull << 31) && !force_idx64())
                dda_kernel<STATS, STRICT, GridCursor<false>><<<grid, block, 0, stream>>>(p);
            else
                dda_kernel<STATS, STRICT, GridCursor<true>><<<grid, block, 0, stream>>>(p);
            break;
        case 1: svo_naive_kernel<STATS, STRICT><<<grid, block, 0, stream>>>(p); break;
        case 2:
            if (p.pool && ray_pool_mode() && !deep) {
                // persistent grid: one wave of resident blocks drawing from the pool
                static int per_sm = 0, sms = 0;
                if (per_sm == 0) {
                    int dev = 0;
                    cudaGetDevice(&dev);
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, esvo_kernel<false, false, 12, true>, BLOCK_THREADS, 0);
                    if (per_sm <= 0) per_sm = 1;
                }
                FrameParams q = p;
                q.pool_blocks_x = grid.x;
                q.pool_total = grid.x * grid.y * (uint32_t)BLOCK_THREADS;
                cudaError_t e = cudaMemsetAsync(q.pool, 0, sizeof(uint32_t), stream);
                if (e != cudaSuccess) return e;
                const uint32_t blocks = min((uint32_t)(per_sm * sms), grid.x * grid.y);
                esvo_kernel<STATS, STRICT, 12, true><<<blocks, block, 0, stream>>>(q);
                break;
            }
            if (deep) esvo_kernel<STATS, STRICT, 24><<<grid, block, 0, stream>>>(p);
            else if (!STATS && !STRICT && p.brick_base != 0xFFFFFFFFu)
                esvo_kernel<false, false, 12, false, true><<<grid, block, 0, stream>>>(p);
            else esvo_kernel<STATS, STRICT, 12><<<grid, block, 0, stream>>>(p);
            break;
        case 3:
            if (deep) svo_df_kernel<STATS, STRICT, 24><<<grid, block, 0, stream>>>(p);
            else svo_df_kernel<STATS, STRICT, 12><<<grid, block, 0, stream>>>(p);
            break;
        case 4:
            if (p.rnodes) svo_rope32_kernel<STATS, STRICT><<<grid, block, 0, stream>>>(p);
            else svo_rope_kernel<STATS, STRICT><<<grid, block, 0, stream>>>(p);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_traversal(int traversal, const FrameParams& p, bool stats, bool strict, cudaStream_t stream) {
    if (p.out_w == 0 || p.out_h == 0) return cudaSuccess;
    if (stats) return strict ? launch_t<true, true>(traversal, p, stream) : launch_t<true, false>(traversal, p, stream);
    return strict ? launch_t<false, true>(traversal, p, stream) : launch_t<false, false>(traversal, p, stream);
}

} // namespace xn
