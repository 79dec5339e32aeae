// xn_util_kernels.cu -- volume re-layout, synthetic-volume and statistics kernels (sm_100a).
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

#include "xn_brick.h"
#include "xn_device.cuh"
#include "xn_kernels.h"
#include "xn_synth.h"

namespace xn {

// ---------------------------------------------------------------------------------
// volume re-layout: 40-byte file nodes -> 64-byte device nodes (one thread per node)
// ---------------------------------------------------------------------------------
__global__ void relayout_nodes_kernel(const uint32_t* __restrict__ raw, uint64_t count, DNode* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t* n = raw + i * 10u;
    DNode d;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t child = n[c];
        if (child >= count) child = 0; // malformed file: never index out of bounds
        const uint32_t* cn = raw + (uint64_t)child * 10u;
        d.slot[c] = make_uint2(child, make_meta(cn[8], cn[9]));
    }
    uint4* o = reinterpret_cast<uint4*>(out + i);
    o[0] = make_uint4(d.slot[0].x, d.slot[0].y, d.slot[1].x, d.slot[1].y);
    o[1] = make_uint4(d.slot[2].x, d.slot[2].y, d.slot[3].x, d.slot[3].y);
    o[2] = make_uint4(d.slot[4].x, d.slot[4].y, d.slot[5].x, d.slot[5].y);
    o[3] = make_uint4(d.slot[6].x, d.slot[6].y, d.slot[7].x, d.slot[7].y);
}

// deepest node of the file (stack sizing, residency limits)
__global__ void max_depth_kernel(const uint32_t* __restrict__ raw, uint64_t count, uint32_t* __restrict__ max_depth) {
    uint32_t my_depth = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        my_depth = max(my_depth, raw[i * 10u + 9u] & 0x7FFFFFFFu);
    for (int o = 16; o > 0; o >>= 1) my_depth = max(my_depth, __shfl_xor_sync(0xFFFFFFFFu, my_depth, o));
    if ((threadIdx.x & 31) == 0 && my_depth) atomicMax(max_depth, my_depth);
}

cudaError_t launch_max_depth(const void* raw40, uint64_t count, uint32_t* d_max_depth, cudaStream_t stream) {
    max_depth_kernel<<<148 * 8, 256, 0, stream>>>((const uint32_t*)raw40, count, d_max_depth);
    return cudaGetLastError();
}

// 40-byte file nodes -> 32-byte rope records (RNode, xn_device.cuh); one thread per node
__global__ void relayout_rnodes_kernel(const uint32_t* __restrict__ raw, uint64_t count, RNode* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t* n = raw + i * 10u;
    uint32_t w[8];
    if (n[9] & 0x80000000u) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            uint32_t t = n[k];
            if (t >= count) t = 0; // malformed file: never index out of bounds
            w[k] = t == 0u ? 0u : (t | ((raw[(uint64_t)t * 10u + 9u] & 0xFu) << 28));
        }
        w[6] = (n[8] & 0x00FFFFFFu) | ((n[9] & 0xFFu) << 24);
        w[7] = RNODE_LEAF_TAG;
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t child = n[c];
            if (child >= count) child = 0;
            w[c] = child | (raw[(uint64_t)child * 10u + 9u] & 0x80000000u);
        }
    }
    uint4* o = reinterpret_cast<uint4*>(out + i);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

cudaError_t launch_relayout_rnodes(const void* raw40, uint64_t count, RNode* out, cudaStream_t stream) {
    const int threads = 256;
    const uint64_t blocks = (count + threads - 1) / threads;
    relayout_rnodes_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const uint32_t*)raw40, count, out);
    return cudaGetLastError();
}

cudaError_t launch_relayout(const void* raw40, uint64_t count, DNode* out, cudaStream_t stream) {
    const int threads = 256;
    const uint64_t blocks = (count + threads - 1) / threads;
    relayout_nodes_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const uint32_t*)raw40, count, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// compact octree residency (CNode, xn_device.cuh): internal nodes only, level order.
//   1. key = depth for internal nodes (root forced to 0), 30 for leaf bricks (internal nodes whose
//      eight children are all leaves), 31 for leaves; stable radix sort of the node indices by key
//      ->  level order with file order inside a level, then the leaf bricks, leaves last: a child
//      word at or above *brick_base_out (a word offset) names a leaf brick, no flag bit needed;
//   2. rank[file index] = position in that order (the compact index);
//   3. one thread per internal node writes its eight child words.
// ---------------------------------------------------------------------------------
__global__ void compact_keys_kernel(const uint32_t* __restrict__ raw, uint64_t count, uint8_t* __restrict__ keys,
                                    uint32_t* __restrict__ vals, unsigned long long* __restrict__ n_internal) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned internal = 0, brick = 0;
    if (i < count) {
        const uint32_t ld = raw[i * 10u + 9u];
        const bool leaf = (ld & 0x80000000u) != 0u;
        internal = (i == 0 || !leaf) ? 1u : 0u; // the root always has a record (a leaf root points at itself)
        if (internal && i != 0) {
            // leaf brick: all eight children are leaves (their colours are this record's eight words)
            brick = 1u;
            for (int c = 0; c < 8; ++c) {
                const uint32_t child = raw[i * 10u + (uint32_t)c];
                if (child >= count || !(raw[(uint64_t)child * 10u + 9u] & 0x80000000u)) brick = 0u;
            }
        }
        keys[i] = i == 0 ? 0 : (leaf ? 31 : (brick ? 30 : (uint8_t)min(ld & 0x7FFFFFFFu, 29u)));
        vals[i] = (uint32_t)i;
    }
    const unsigned n = __popc(__ballot_sync(0xFFFFFFFFu, internal != 0u));
    const unsigned nb = __popc(__ballot_sync(0xFFFFFFFFu, brick != 0u));
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(n_internal, (unsigned long long)n);
    if ((threadIdx.x & 31) == 0 && nb) atomicAdd(n_internal + 1, (unsigned long long)nb);
}

__global__ void compact_rank_kernel(const uint32_t* __restrict__ order, uint64_t n_internal, uint32_t* __restrict__ rank) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_internal) rank[order[j]] = (uint32_t)j;
}

__global__ void compact_emit_kernel(const uint32_t* __restrict__ raw, uint64_t count, const uint32_t* __restrict__ order,
                                    const uint32_t* __restrict__ rank, uint64_t n_internal, CNode* __restrict__ out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_internal) return;
    const uint32_t* n = raw + (uint64_t)order[j] * 10u;
    uint32_t w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t child = n[c];
        if (child >= count) child = 0; // malformed file: never index out of bounds
        const uint32_t* cn = raw + (uint64_t)child * 10u;
        w[c] = (cn[9] & 0x80000000u) ? make_meta(cn[8], cn[9]) : rank[child] << 3; // word offset of the record
    }
    uint4* o = reinterpret_cast<uint4*>(out + j);
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

cudaError_t build_compact_nodes(const void* raw40, uint64_t count, CNode** out, uint64_t* n_internal_out,
                                uint32_t* brick_base_out, cudaStream_t stream) {
    *out = nullptr;
    uint8_t *k_in = nullptr, *k_out = nullptr;
    uint32_t *v_in = nullptr, *v_out = nullptr, *rank = nullptr;
    unsigned long long* d_n = nullptr;
    void* tmp = nullptr;
    CNode* nodes = nullptr;
    cudaError_t e = cudaSuccess;
    auto done = [&](cudaError_t err) {
        cudaFree(k_in), cudaFree(k_out), cudaFree(v_in), cudaFree(v_out), cudaFree(rank), cudaFree(d_n), cudaFree(tmp);
        if (err != cudaSuccess && nodes) cudaFree(nodes);
        return err;
    };
#define XN_TRY(call) if ((e = (call)) != cudaSuccess) return done(e)
    XN_TRY(cudaMalloc(&k_in, count));
    XN_TRY(cudaMalloc(&k_out, count));
    XN_TRY(cudaMalloc(&v_in, count * 4));
    XN_TRY(cudaMalloc(&v_out, count * 4));
    XN_TRY(cudaMalloc(&rank, count * 4));
    XN_TRY(cudaMalloc(&d_n, 16));
    XN_TRY(cudaMemsetAsync(d_n, 0, 16, stream));
    const int threads = 256;
    const unsigned blocks = (unsigned)((count + threads - 1) / threads);
    compact_keys_kernel<<<blocks, threads, 0, stream>>>((const uint32_t*)raw40, count, k_in, v_in, d_n);
    XN_TRY(cudaGetLastError());
    size_t tmp_bytes = 0;
    XN_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)count, 0, 5, stream));
    XN_TRY(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    XN_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int64_t)count, 0, 5, stream));
    unsigned long long counts[2] = {0, 0}; // internal nodes, of which leaf bricks
    XN_TRY(cudaMemcpyAsync(counts, d_n, 16, cudaMemcpyDeviceToHost, stream));
    XN_TRY(cudaStreamSynchronize(stream));
    const unsigned long long n_internal = counts[0];
    if (n_internal > (1ull << 28)) return done(cudaErrorInvalidValue); // word offsets must fit 31 bits
    XN_TRY(cudaMalloc(&nodes, n_internal * sizeof(CNode)));
    const unsigned iblocks = (unsigned)((n_internal + threads - 1) / threads);
    compact_rank_kernel<<<iblocks, threads, 0, stream>>>(v_out, n_internal, rank);
    XN_TRY(cudaGetLastError());
    compact_emit_kernel<<<iblocks, threads, 0, stream>>>((const uint32_t*)raw40, count, v_out, rank, n_internal, nodes);
    XN_TRY(cudaGetLastError());
    XN_TRY(cudaStreamSynchronize(stream));
#undef XN_TRY
    *out = nodes;
    *n_internal_out = n_internal;
    *brick_base_out = (uint32_t)((n_internal - counts[1]) << 3);
    return done(cudaSuccess);
}

// ---------------------------------------------------------------------------------
// svo_naive entry table: one thread per cell of the 2^levels grid walks find()'s first `levels`
// levels (the child at level l is bit l of each cell coordinate) and records where it ends: a
// leaf word, or the word offset of the internal node at depth `levels`.
// ---------------------------------------------------------------------------------
__global__ void top_table_kernel(const CNode* __restrict__ nodes, uint32_t root_meta, uint32_t levels,
                                 uint32_t* __restrict__ table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << (3 * levels))) return;
    const uint32_t m = (1u << levels) - 1u;
    const uint32_t cz = i & m, cy = (i >> levels) & m, cx = i >> (2 * levels);
    uint32_t w = (root_meta & META_LEAF) ? root_meta : 0u; // leaf root: its meta; else word offset 0
    for (uint32_t l = 0; l < levels && !(w & META_LEAF); ++l) {
        const uint32_t sh = levels - 1u - l;
        const uint32_t child = (((cx >> sh) & 1u) << 2) | (((cy >> sh) & 1u) << 1) | ((cz >> sh) & 1u);
        w = reinterpret_cast<const uint32_t*>(nodes)[w | child];
    }
    table[i] = w;
}

cudaError_t launch_top_table(const CNode* nodes, uint32_t root_meta, uint32_t levels, uint32_t* table,
                             cudaStream_t stream) {
    const uint32_t n = 1u << (3 * levels);
    top_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(nodes, root_meta, levels, table);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// grid re-layout: x-major linear <-> bricked (xn_brick.h).  One thread per bricked slot, so the
// bricked side is accessed in order; the linear side is touched in 2x2x2 / 4x4x2 groups that
// stay inside a few cache lines.  Padding slots are written as 0 (= the border colour).
// ---------------------------------------------------------------------------------
__global__ void brick_grid_kernel(const uint32_t* __restrict__ linear, uint32_t* __restrict__ bricked, BrickLayout L,
                                  uint32_t nx, uint32_t ny, uint32_t nz) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = brick_coord(L, 0, i), y = brick_coord(L, 1, i), z = brick_coord(L, 2, i);
        uint32_t v = 0;
        if (x < nx && y < ny && z < nz) v = linear[(uint64_t)x + (uint64_t)y * nx + (uint64_t)z * nx * ny];
        bricked[i] = v;
    }
}

__global__ void unbrick_grid_kernel(const uint32_t* __restrict__ bricked, uint32_t* __restrict__ linear, BrickLayout L,
                                    uint32_t nx, uint32_t ny, uint32_t nz) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = brick_coord(L, 0, i), y = brick_coord(L, 1, i), z = brick_coord(L, 2, i);
        if (x < nx && y < ny && z < nz) linear[(uint64_t)x + (uint64_t)y * nx + (uint64_t)z * nx * ny] = bricked[i];
    }
}

cudaError_t launch_brick_grid(const uint32_t* linear, uint32_t* bricked, const BrickLayout& L, uint32_t nx, uint32_t ny,
                              uint32_t nz, cudaStream_t stream) {
    brick_grid_kernel<<<148 * 16, 256, 0, stream>>>(linear, bricked, L, nx, ny, nz);
    return cudaGetLastError();
}

cudaError_t launch_unbrick_grid(const uint32_t* bricked, uint32_t* linear, const BrickLayout& L, uint32_t nx,
                                uint32_t ny, uint32_t nz, cudaStream_t stream) {
    unbrick_grid_kernel<<<148 * 16, 256, 0, stream>>>(bricked, linear, L, nx, ny, nz);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// volume ingest: raw TIFF sample bytes of one z slice (rows in file order) -> RGBA8 voxels of
// the grid slice.  What TIFFReadRGBAImage does on the host in the reference
// (src/model/Grid.cpp:60-75): grey / RGB expansion, white-is-zero inversion, pre-multiplication
// of unassociated alpha ((c * a + 127) / 255), bottom-up raster order.  One thread per pixel.
// ---------------------------------------------------------------------------------
__global__ void tiff_decode_kernel(const uint8_t* __restrict__ raw, uint32_t* __restrict__ slice, uint32_t W, uint32_t H,
                                   TiffDecode f) {
    const uint64_t n = (uint64_t)W * H;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % W), row = (uint32_t)(i / W);
        uint32_t r, g, b, a = 255u;
        if (f.samples == 4u && f.photometric == 2u) {
            const uint32_t v = reinterpret_cast<const uint32_t*>(raw)[i]; // staging is 256-byte aligned
            r = v & 0xFFu, g = (v >> 8) & 0xFFu, b = (v >> 16) & 0xFFu, a = v >> 24;
        } else {
            const uint8_t* px = raw + i * f.samples;
            uint32_t cs;
            if (f.photometric == 2u) {
                r = px[0], g = px[1], b = px[2];
                cs = 3;
            } else {
                r = g = b = f.photometric == 0u ? 255u - px[0] : px[0];
                cs = 1;
            }
            if (f.has_alpha) a = px[cs];
        }
        if (f.unassociated) {
            r = (r * a + 127u) / 255u;
            g = (g * a + 127u) / 255u;
            b = (b * a + 127u) / 255u;
        }
        const uint32_t y = (f.flip & 1u) ? H - 1u - row : row;   // Orientation tag: rows reversed
        const uint32_t xo = (f.flip & 2u) ? W - 1u - x : x;      // ... columns reversed
        slice[(uint64_t)y * W + xo] = r | (g << 8) | (b << 16) | (a << 24);
    }
}

cudaError_t launch_tiff_decode(const uint8_t* raw, uint32_t* slice, uint32_t W, uint32_t H, const TiffDecode& f,
                               cudaStream_t stream) {
    const uint64_t n = (uint64_t)W * H;
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
    tiff_decode_kernel<<<blocks, 256, 0, stream>>>(raw, slice, W, H, f);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// synthetic volumes (bit-identical to the host generator in xn_synth.h)
// ---------------------------------------------------------------------------------
__global__ void synth_kernel(uint32_t* __restrict__ grid, SynthSpec spec) {
    const uint64_t n = (uint64_t)spec.nx * spec.ny * spec.nz;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % spec.nx);
        const uint32_t y = (uint32_t)((i / spec.nx) % spec.ny);
        const uint32_t z = (uint32_t)(i / ((uint64_t)spec.nx * spec.ny));
        grid[i] = synth_voxel(spec, x, y, z);
    }
}

cudaError_t launch_synth(uint32_t* grid, const SynthSpec& spec, cudaStream_t stream) {
    synth_kernel<<<148 * 16, 256, 0, stream>>>(grid, spec);
    return cudaGetLastError();
}

// share of black (r = g = b = 0) voxels, from every `stride`-th voxel
__global__ void count_black_kernel(const uint32_t* __restrict__ grid, uint64_t n, uint64_t stride,
                                   unsigned long long* __restrict__ out) {
    unsigned long long black = 0, seen = 0;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * stride; i < n;
         i += (uint64_t)gridDim.x * blockDim.x * stride) {
        black += (grid[i] & 0x00FFFFFFu) == 0u;
        ++seen;
    }
    for (int o = 16; o > 0; o >>= 1) {
        black += __shfl_xor_sync(0xFFFFFFFFu, black, o);
        seen += __shfl_xor_sync(0xFFFFFFFFu, seen, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], black);
        atomicAdd(&out[1], seen);
    }
}

cudaError_t launch_count_black(const uint32_t* grid, uint64_t n, uint64_t stride, unsigned long long* out,
                               cudaStream_t stream) {
    count_black_kernel<<<148 * 8, 256, 0, stream>>>(grid, n, stride, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// DDA skip table (xn_device.cuh, FrameParams::skip_table): which B^3 bricks of the grid hold one
// colour, and how far that colour extends in each of the eight directions of travel.  The DDA's
// fetch (resources/dda.comp:45) is the only thing the table replaces: a ray still takes every
// step of the reference's march, but texels the table lets it know are not read.
//   entry (16 bytes) = { x: rgb | uniform << 24, y: 0, z: radius bytes of octants 0-3, w: 4-7 }
//   octant o = (dir.x > 0) | (dir.y > 0) << 1 | (dir.z > 0) << 2.
//   radius k of octant o (0 for a mixed brick): the k^3 bricks b + s * [0, k)^3, s = the octant's
//   signs, all hold this one colour -- so a ray travelling in octant o from any voxel of the brick
//   finds this colour for R_i = r_i + (k - 1) B voxels on axis i, r_i = voxels left to the brick
//   face.  (Largest-cube dynamic programme: k(b) = 1 + min over the seven bricks b + s * delta.)
// The table has a one-brick border on every side standing for the sampler's border colour
// (transparent black, src/render/DdaRaytraceAlgorithm.cpp:26-29); voxels of an edge brick that
// lie outside the grid count as black too.  Alpha is ignored: the march reads texel.rgb only.
// ---------------------------------------------------------------------------------
constexpr uint32_t SKIP_UNIFORM = 1u << 24;

__global__ void skip_classify_kernel(const uint32_t* __restrict__ grid, uint32_t nx, uint32_t ny, uint32_t nz,
                                     uint32_t shift, uint32_t tx, uint32_t ty, uint4* __restrict__ table) {
    // one block per brick; thread = one x row of the brick
    const uint32_t B = 1u << shift;
    const uint32_t bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    const uint32_t x0 = bx << shift, y0 = by << shift, z0 = bz << shift;
    const uint32_t ref = grid[(uint64_t)x0 + (uint64_t)y0 * nx + (uint64_t)z0 * nx * ny] & 0x00FFFFFFu;
    bool same = true;
    for (uint32_t row = threadIdx.x; row < B * B; row += blockDim.x) {
        const uint32_t y = y0 + (row & (B - 1u)), z = z0 + (row >> shift);
        if (y < ny && z < nz) {
            const uint32_t* r = grid + ((uint64_t)y * nx + (uint64_t)z * nx * ny);
            for (uint32_t x = x0; x < x0 + B; ++x) same &= ((x < nx ? r[x] : 0u) & 0x00FFFFFFu) == ref;
        } else {
            same &= ref == 0u;
        }
    }
    same = __syncthreads_and(same) != 0;
    if (threadIdx.x == 0)
        table[((uint64_t)(bz + 1u) * ty + (by + 1u)) * tx + (bx + 1u)] =
            same ? make_uint4(ref | SKIP_UNIFORM, 0u, 0x01010101u, 0x01010101u) : make_uint4(ref, 0u, 0u, 0u);
}

__global__ void skip_fill_kernel(uint4* __restrict__ table, uint64_t n, uint4 value) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        table[i] = value;
}

// one relaxation step of the eight radii: k' = 1 + min over the seven bricks one step along the
// octant of (same colour ? their k : 0), capped; beyond the table everything is border colour
// with unbounded radius.  Starting from k = 1, iteration n yields min(true k, n + 1).
__global__ void skip_relax_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t tx, uint32_t ty,
                                  uint32_t tz, uint32_t cap) {
    const uint64_t n = (uint64_t)tx * ty * tz;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 e = in[i];
        if (e.x & SKIP_UNIFORM) {
            const int x = (int)(i % tx), y = (int)((i / tx) % ty), z = (int)(i / ((uint64_t)tx * ty));
            const uint32_t beyond = (e.x & 0x00FFFFFFu) == 0u ? cap : 0u;
            // radii of the 26 neighbours in their own octant directions are read per octant below;
            // a neighbour of another colour (or a mixed one) contributes 0
            uint32_t lo = 0, hi = 0;
            for (uint32_t o = 0; o < 8u; ++o) {
                const int sx = (o & 1u) ? 1 : -1, sy = (o & 2u) ? 1 : -1, sz = (o & 4u) ? 1 : -1;
                uint32_t m = cap;
                for (uint32_t dl = 1; dl < 8u; ++dl) {
                    const int X = x + ((dl & 1u) ? sx : 0), Y = y + ((dl & 2u) ? sy : 0), Z = z + ((dl & 4u) ? sz : 0);
                    uint32_t nd;
                    if (X < 0 || Y < 0 || Z < 0 || X >= (int)tx || Y >= (int)ty || Z >= (int)tz) {
                        nd = beyond;
                    } else {
                        const uint4 ne = in[((uint64_t)Z * ty + Y) * tx + X];
                        nd = ne.x == e.x ? (((o < 4u ? ne.z : ne.w) >> (8u * (o & 3u))) & 0xFFu) : 0u;
                    }
                    m = min(m, nd);
                }
                const uint32_t k = min(cap, m + 1u);
                if (o < 4u) lo |= k << (8u * o);
                else hi |= k << (8u * (o - 4u));
            }
            e.z = lo;
            e.w = hi;
        }
        out[i] = e;
    }
}

__global__ void skip_count_kernel(const uint4* __restrict__ table, uint64_t n, unsigned long long* __restrict__ out) {
    unsigned long long uniform = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        uniform += (table[i].x & SKIP_UNIFORM) != 0u;
    for (int o = 16; o > 0; o >>= 1) uniform += __shfl_xor_sync(0xFFFFFFFFu, uniform, o);
    if ((threadIdx.x & 31) == 0 && uniform) atomicAdd(out, uniform);
}

cudaError_t build_skip_table(const uint32_t* grid, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t shift, uint32_t cap,
                             uint4** table_out, uint32_t dims_out[3], double* uniform_fraction_out, cudaStream_t stream) {
    *table_out = nullptr;
    if (shift < 1u || shift > 5u || cap < 1u || cap > 255u) return cudaErrorInvalidValue;
    // the march's promise covers at most (cap - 1) B + B - 1 steps per axis; the rounding bound of
    // the kernel's t_safe (a factor 1 - 2^-14) holds for up to ~1000 repeated additions
    while (cap > 1u && ((cap << shift) > 1000u)) --cap;
    const uint32_t B = 1u << shift;
    const uint32_t bx = (nx + B - 1u) >> shift, by = (ny + B - 1u) >> shift, bz = (nz + B - 1u) >> shift;
    if (by > 65535u || bz > 65535u) return cudaErrorInvalidValue;
    const uint32_t tx = bx + 2u, ty = by + 2u, tz = bz + 2u;
    const uint64_t n = (uint64_t)tx * ty * tz;
    if (n >= (1ull << 31)) return cudaErrorInvalidValue; // the kernel indexes the table with 32 bits
    uint4 *a = nullptr, *b = nullptr;
    cudaError_t e = cudaMalloc(&a, n * sizeof(uint4));
    if (e == cudaSuccess) e = cudaMalloc(&b, n * sizeof(uint4));
    if (e == cudaSuccess) {
        // border: black, radius 1 so far
        skip_fill_kernel<<<148 * 8, 256, 0, stream>>>(a, n, make_uint4(SKIP_UNIFORM, 0u, 0x01010101u, 0x01010101u));
        skip_classify_kernel<<<dim3(bx, by, bz), B * B < 64u ? B * B : 64u, 0, stream>>>(grid, nx, ny, nz, shift, tx, ty, a);
        for (uint32_t it = 1; it < cap; ++it) {
            skip_relax_kernel<<<148 * 8, 256, 0, stream>>>(a, b, tx, ty, tz, cap);
            uint4* t = a;
            a = b;
            b = t;
        }
        e = cudaGetLastError();
        // share of the grid's own bricks (border layer excluded) that hold one colour
        unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(b);
        unsigned long long cnt = 0;
        if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, 8, stream);
        if (e == cudaSuccess) {
            skip_count_kernel<<<148 * 8, 256, 0, stream>>>(a, n, d_cnt);
            e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        const uint64_t own = (uint64_t)bx * by * bz;
        if (uniform_fraction_out) *uniform_fraction_out = (double)(cnt - (n - own)) / (double)own;
    }
    cudaFree(b);
    if (e != cudaSuccess) {
        cudaFree(a);
        return e;
    }
    *table_out = a;
    dims_out[0] = tx, dims_out[1] = ty, dims_out[2] = tz;
    return cudaSuccess;
}

// sum reduction of the stats arrays (totals for the roofline accounting)
__global__ void stats_totals_kernel(const uint32_t* __restrict__ steps, const unsigned long long* __restrict__ bytes,
                                    uint64_t n, unsigned long long* __restrict__ totals) {
    unsigned long long s = 0, b = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        s += steps[i];
        b += bytes[i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&totals[0], s);
        atomicAdd(&totals[1], b);
    }
}

// touch map -> (bits set = voxels fetched, non-zero bytes = 32-byte sectors of the linear layout fetched)
__global__ void touch_count_kernel(const uint32_t* __restrict__ bits, uint64_t words, unsigned long long* __restrict__ out) {
    unsigned long long voxels = 0, sectors = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = bits[i];
        voxels += (unsigned)__popc(w);
        sectors += ((w & 0xFFu) != 0u) + ((w & 0xFF00u) != 0u) + ((w & 0xFF0000u) != 0u) + ((w & 0xFF000000u) != 0u);
    }
    for (int o = 16; o > 0; o >>= 1) {
        voxels += __shfl_down_sync(0xFFFFFFFFu, voxels, o);
        sectors += __shfl_down_sync(0xFFFFFFFFu, sectors, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (voxels) atomicAdd(out, voxels);
        if (sectors) atomicAdd(out + 1, sectors);
    }
}
cudaError_t launch_touch_count(const uint32_t* bits, uint64_t words, unsigned long long* out2, cudaStream_t stream) {
    const unsigned blocks = (unsigned)std::min<uint64_t>((words + 255) / 256, 148u * 16u);
    touch_count_kernel<<<blocks ? blocks : 1, 256, 0, stream>>>(bits, words, out2);
    return cudaGetLastError();
}

cudaError_t launch_stats_totals(const uint32_t* steps, const unsigned long long* bytes, uint64_t n,
                                unsigned long long* totals, cudaStream_t stream) {
    stats_totals_kernel<<<148 * 4, 256, 0, stream>>>(steps, bytes, n, totals);
    return cudaGetLastError();
}

} // namespace xn
