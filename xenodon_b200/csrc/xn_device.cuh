// xn_device.cuh -- device-side types and binary32 helpers shared by the traversal kernels.
//
// Arithmetic contract: every geometric quantity (ray set-up, slab tests, DDA side
// distances, octree descent decisions) is computed in IEEE binary32 in the order the
// reference shader source writes it, with no FMA contraction (this translation unit is
// compiled with -fmad=false and uses IEEE division / square root), so the sequence of
// voxels / nodes a ray visits is identical to the CPU oracle's.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xn {

// ---------------------------------------------------------------------------------
// Device volume layouts
// ---------------------------------------------------------------------------------

// Octree node as resident in HBM: 8 child descriptors of 8 bytes, 64-byte aligned, so one
// 64-bit load yields everything the shaders read about a child with three dependent 32-bit
// loads (children[i], nodes[child].is_leaf_depth, nodes[child].color; resources/octree.glsl:6-14).
//   slot.x = child node index (for a rope-file leaf, slots 0..5 are the ropes)
//   slot.y = meta of THAT child: bit 31 leaf | depth << 24 | b << 16 | g << 8 | r
struct __align__(64) DNode {
    uint2 slot[8];
};
static_assert(sizeof(DNode) == 64, "device node must be 64 bytes");

// Compact octree residency read by svo_naive / svo_df / esvo: INTERNAL nodes only, 32 bytes
// (one sector) each, numbered level by level (root = 0, then depth 1, ...; file order inside a
// level), so the upper levels are one contiguous range that an L2 access-policy window keeps
// resident.  One 32-bit word per child:
//   leaf child     : bit 31 set | depth << 24 | b << 16 | g << 8 | r   (the same bits as `meta`)
//   internal child : the word offset of its record = compact index * 8 (bit 31 clear; up to 2^28
//                    internal nodes, i.e. trees of about 2^31 nodes)
// A leaf's own node record is never read by these three traversals (everything they need about a
// leaf is in its parent's word), so leaves -- 7/8 of a tree -- occupy no space here.
struct __align__(32) CNode {
    uint32_t w[8];
};
static_assert(sizeof(CNode) == 32, "compact node must be 32 bytes");

// Rope-tree residency read by svo_rope when the tree allows it (fewer than 2^28 nodes, depth <= 15):
// every node of the file, 32 bytes (one sector) each, in file order.
//   internal node : w[c] = child index | child_is_leaf << 31   (a child's depth is its parent's + 1)
//   leaf node     : w[0..5] = the six ropes, neighbour index | neighbour depth << 28 (0 = no
//                   neighbour, as in the file); w[6] = rgb | own depth << 24; w[7] = RNODE_LEAF_TAG
// A ray leaving a leaf needs that leaf's colour and ONE of its ropes: both sit in the one sector
// the visit fetches, where the 64-byte DNode spreads the six (rope, neighbour meta) pairs over two
// sectors that different rays touch -- half the DRAM traffic per touched leaf.
struct __align__(32) RNode {
    uint32_t w[8];
};
static_assert(sizeof(RNode) == 32, "rope node must be 32 bytes");
constexpr uint32_t RNODE_LEAF_TAG = 0xFFFFFFFFu;
constexpr uint64_t RNODE_MAX_NODES = 1ull << 28;
constexpr uint32_t RNODE_MAX_DEPTH = 15;

// svo_naive entry table: at most this many levels of the tree are folded into it (8: 64 MiB)
#ifndef XN_TOP_LEVELS_MAX
#define XN_TOP_LEVELS_MAX 8
#endif
constexpr uint32_t TOP_LEVELS_MAX = XN_TOP_LEVELS_MAX;
constexpr uint32_t META_LEAF = 0x80000000u;
__host__ __device__ inline uint32_t make_meta(uint32_t color, uint32_t is_leaf_depth) {
    return (is_leaf_depth & META_LEAF) | ((is_leaf_depth & 0x1Fu) << 24) | (color & 0x00FFFFFFu);
}
__device__ __forceinline__ bool meta_is_leaf(uint32_t m) { return (m & META_LEAF) != 0u; }
__device__ __forceinline__ uint32_t meta_depth(uint32_t m) { return (m >> 24) & 0x1Fu; }

// ---------------------------------------------------------------------------------
// Per-frame kernel parameters (push constants + uniform buffer of the reference,
// resources/common.glsl:6-35), passed by value as a __grid_constant__.
// ---------------------------------------------------------------------------------
struct FrameParams {
    float fwd[3], up[3], pos[3]; // pos = translation / voxel_ratio (src/render/Renderer.cpp:62)
    int32_t out_x, out_y;
    uint32_t out_w, out_h;
    int32_t disp_x, disp_y;
    uint32_t disp_w, disp_h;
    float ratio[3];
    uint32_t model_dim[3];
    float emission;

    uint32_t* target;      // RGBA8 pixels, may point into a peer device's frame
    uint64_t target_stride; // in pixels

    const uint32_t* grid;   // RGBA8 voxels: x fastest, or bricked (xn_brick.h) when bk_slots != 0
    uint32_t nx, ny, nz;
    // bricked residency: index bits owned by x, y, z (low 32 bits), position of each axis'
    // brick-index field, the axis on top, z's full 64-bit mask, number of voxel slots (0 = linear)
    uint32_t bk_mask[3], bk_hs[3], bk_top;
    uint64_t bk_mask_z64, bk_slots;
    // texture residency: the grid as a 3-D CUDA array (block-linear tiling, border = 0) seen through
    // two texture objects: channels as c / 255 floats (fast mode) and as raw bytes (strict mode)
    unsigned long long tex_unorm, tex_raw;
    // DDA skip table (built at upload, xn_util_kernels.cu): one 16-byte entry per 2^skip_shift-voxel
    // brick = { rgb | uniform << 24, 0, radii of octants 0-3, radii of octants 4-7 }: from a uniform
    // brick, the k^3 bricks in the direction of travel hold this colour; extent skip_dim[] bricks
    // including a one-brick border of border colour.  nullptr = no table (every texel is fetched)
    const uint4* skip_table;
    uint32_t skip_dim[3], skip_shift;
    const DNode* nodes;   // svo_rope (and the file-order view of the tree)
    const RNode* rnodes;  // svo_rope, 32-byte records (nullptr: the tree exceeds their limits, nodes is read)
    const CNode* cnodes;  // svo_naive, svo_df, esvo
    // child words at or above this word offset name leaf bricks: records whose eight children are all
    // leaves (sorted behind the other internal nodes); 0xFFFFFFFF = none / not used
    uint32_t brick_base;
    // svo_naive: what find() reaches after its first top_levels levels, for each of the
    // 2^(3 top_levels) aligned cells of the cube (a leaf word, or the word offset of an internal
    // node), index = cx << 2 top_levels | cy << top_levels | cz; top_levels = min(tree depth, 8)
    const uint32_t* top_table;
    uint32_t top_levels;
    uint32_t root_meta;
    uint32_t max_depth;     // deepest node depth in the tree (stack sizing)

    // row interleave: this launch shades only the 16-row stripes s with s % il_count == il_index
    // (the same partition as il_count*... `device {}` blocks of 16 rows each, in one launch)
    uint32_t il_count, il_index;
    // grid volumes with a substantial share of black voxels (decided at upload): the DDA skips
    // the colour arithmetic of warp-wide empty stretches
    uint32_t skip_empty;

    // persistent ray pool (esvo_kernel<..., POOL>): counter of rays handed out, rays of this launch,
    // 16-pixel block columns of the frame; nullptr = static one-thread-per-pixel launch
    uint32_t* pool;
    uint32_t pool_total, pool_blocks_x;

    uint32_t* touch_bits;           // touch pass only: one bit per voxel (x-major linear order) fetched
    uint32_t* steps_out;            // stats pass only
    unsigned long long* bytes_out;  // stats pass only
};

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 F3(float x, float y, float z) { return f3{x, y, z}; }

// GLSL min/max: min(x, y) = y < x ? y : x.  Identical to fminf/fmaxf for the finite,
// non-signed-zero-sensitive values on this path; the select form is used where a zero of
// either sign can appear so results stay bit-identical to the oracle.
__device__ __forceinline__ float gmin(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float gmax(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float min_elem(f3 v) { return gmin(v.x, gmin(v.y, v.z)); }
__device__ __forceinline__ float max_elem(f3 v) { return gmax(v.x, gmax(v.y, v.z)); }
__device__ __forceinline__ float gsign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ float gmod(float x, float y) { return x - y * floorf(x / y); }
// mod(x, y) for y an exact power of two (node sizes 2^-depth): x / y equals x * (1 / y) bit for
// bit because scaling by a power of two is exact, so the IEEE division (a reciprocal on the
// quarter-rate pipe plus fix-up steps) becomes one multiply.  ry = pow2_reciprocal(y).
__device__ __forceinline__ float pow2_reciprocal(float y) { return __int_as_float(0x7F000000 - __float_as_int(y)); }
__device__ __forceinline__ float gmod_pow2(float x, float y, float ry) { return x - y * floorf(x * ry); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return F3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ f3 normalize3(f3 a) {
    float len = sqrtf(dot3(a, a));
    return F3(a.x / len, a.y / len, a.z / len);
}

// exact (float)b / 255.0f for b in 0..255 without an IEEE division: one multiply by the
// rounded reciprocal plus one FMA Newton correction (verified for all 256 inputs in
// tests/test_host_logic.py::test_unorm8_reciprocal_identity).
__device__ __forceinline__ float unorm8(uint32_t b) {
    const float r = 1.0f / 255.0f;
    float x = (float)b;
    float q = x * r;
    float rem = __fmaf_rn(-q, 255.0f, x);
    return __fmaf_rn(rem, r, q);
}

// imageStore to rgba8: clamp to [0,1], scale by 255, round half to even; NaN -> 0
__device__ __forceinline__ uint32_t pack_unorm8(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v > 1.0f) v = 1.0f;
    return (uint32_t)__float2uint_rn(v * 255.0f);
}

// resources/common.glsl:40-56 -- camera ray for a pixel, in the reference's operation order
__device__ __forceinline__ f3 make_ray(const FrameParams& p, int32_t pixel_x, int32_t pixel_y) {
    float uvx = (float)(pixel_x - p.disp_x) / (float)p.disp_w;
    float uvy = (float)(pixel_y - p.disp_y) / (float)p.disp_h;
    uvx -= 0.5f;
    uvy -= 0.5f;
    uvy *= (float)p.disp_h / (float)p.disp_w;

    f3 dir = F3(p.fwd[0], p.fwd[1], p.fwd[2]);
    f3 up = F3(p.up[0], p.up[1], p.up[2]);
    f3 right = normalize3(cross3(up, dir));
    up = normalize3(cross3(right, dir));

    f3 rd = F3((uvx * right.x + uvy * up.x) + dir.x, (uvx * right.y + uvy * up.y) + dir.y,
               (uvx * right.z + uvy * up.z) + dir.z);
    rd = normalize3(rd);
    rd = normalize3(F3(rd.x / p.ratio[0], rd.y / p.ratio[1], rd.z / p.ratio[2]));

    const float epsilon = 1.1920928955078125e-07f; // exp2(-23)
    if (fabsf(rd.x) < epsilon) rd.x = epsilon;
    if (fabsf(rd.y) < epsilon) rd.y = epsilon;
    if (fabsf(rd.z) < epsilon) rd.z = epsilon;
    return rd;
}

// resources/common.glsl:68-72
__device__ __forceinline__ float voxel_emission_coeff(const FrameParams& p, f3 rd) {
    f3 rd2 = F3(rd.x * rd.x, rd.y * rd.y, rd.z * rd.z);
    f3 dim2 = F3(p.ratio[0] * p.ratio[0], p.ratio[1] * p.ratio[1], p.ratio[2] * p.ratio[2]);
    return p.emission * sqrtf(dot3(rd2, dim2) / dot3(rd2, F3(1.f, 1.f, 1.f)));
}

__device__ __forceinline__ uint32_t pack_pixel(f3 c) {
    return pack_unorm8(c.x) | (pack_unorm8(c.y) << 8) | (pack_unorm8(c.z) << 16) | 0xFF000000u;
}

// Pixel owned by this thread: each warp shades a TILE_W x TILE_H pixel tile (32 pixels; rays
// of a warp stay spatially coherent); a block is BLOCK_WARPS_X tiles wide and 16 rows tall (the
// stripe height of xn_set_interleave).  Tunables: XN_TILE_W in {4, 8, 16, 32}, XN_BLOCK_WARPS_X.
#ifndef XN_TILE_W
#define XN_TILE_W 8
#endif
#ifndef XN_TILE_MORTON
#define XN_TILE_MORTON 1
#endif
#ifndef XN_BLOCK_WARPS_X
#define XN_BLOCK_WARPS_X 2
#endif
constexpr int TILE_W = XN_TILE_W, TILE_H = 32 / TILE_W;
constexpr int BLOCK_WARPS_X = XN_BLOCK_WARPS_X, BLOCK_WARPS_Y = 16 / TILE_H;
constexpr int BLOCK_THREADS = 32 * BLOCK_WARPS_X * BLOCK_WARPS_Y;
constexpr int BLOCK_W = TILE_W * BLOCK_WARPS_X, BLOCK_H = 16;
static_assert(TILE_W * TILE_H == 32 && BLOCK_WARPS_Y * TILE_H == 16, "tile must hold one warp");
// MORTON: lanes in Morton order inside the 8x4 tile (lane bits x0 y0 x1 y1 x2), so 4 consecutive
// lanes -- the quad the texture unit works on -- are a 2x2 pixel block and 16 lanes a 4x4 block.
// Used by the texture-path DDA (cfg3 +15 %, cfg4 +23 %); the LDG kernels coalesce over the whole
// warp and are indifferent (DDA) or slightly worse (ESVO -1.8 %) with it.
template <bool MORTON = false>
__device__ __forceinline__ void thread_pixel(const FrameParams& p, uint32_t& ix, uint32_t& iy) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t lx = lane % TILE_W, ly = lane / TILE_W;
    if (MORTON && XN_TILE_MORTON && TILE_W == 8) {
        lx = (lane & 1u) | ((lane >> 1) & 2u) | ((lane >> 2) & 4u);
        ly = ((lane >> 1) & 1u) | ((lane >> 2) & 2u);
    }
    ix = blockIdx.x * BLOCK_W + (warp % BLOCK_WARPS_X) * TILE_W + lx;
    iy = (blockIdx.y * p.il_count + p.il_index) * BLOCK_H + (warp / BLOCK_WARPS_X) * TILE_H + ly;
}

} // namespace xn
