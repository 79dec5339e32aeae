// xn_kernels.h -- host-callable launchers of the device code in xn_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include "xn_brick.h"
#include "xn_device.cuh"
#include "xn_synth.h"

namespace xn {
cudaError_t launch_traversal(int traversal, const FrameParams& p, bool stats, bool strict, cudaStream_t stream);
cudaError_t launch_relayout(const void* raw40, uint64_t count, DNode* out, cudaStream_t stream);
cudaError_t launch_max_depth(const void* raw40, uint64_t count, uint32_t* d_max_depth, cudaStream_t stream);
// 32-byte rope records (RNode): the caller checks RNODE_MAX_NODES / RNODE_MAX_DEPTH first
cudaError_t launch_relayout_rnodes(const void* raw40, uint64_t count, RNode* out, cudaStream_t stream);
// compact residency of the octree (internal nodes only, level order); *out is cudaMalloc'ed
cudaError_t build_compact_nodes(const void* raw40, uint64_t count, CNode** out, uint64_t* n_internal_out,
                                uint32_t* brick_base_out,
                                cudaStream_t stream);
// svo_naive entry table (2^(3 levels) words)
cudaError_t launch_top_table(const CNode* nodes, uint32_t root_meta, uint32_t levels, uint32_t* table,
                             cudaStream_t stream);
cudaError_t launch_brick_grid(const uint32_t* linear, uint32_t* bricked, const BrickLayout& L, uint32_t nx, uint32_t ny,
                              uint32_t nz, cudaStream_t stream);
cudaError_t launch_unbrick_grid(const uint32_t* bricked, uint32_t* linear, const BrickLayout& L, uint32_t nx,
                                uint32_t ny, uint32_t nz, cudaStream_t stream);
// raw TIFF samples of one slice -> RGBA8 grid slice (fields as xn::TiffSliceFormat)
struct TiffDecode {
    uint32_t samples, photometric, has_alpha, unassociated, flip;
};
cudaError_t launch_tiff_decode(const uint8_t* raw, uint32_t* slice, uint32_t W, uint32_t H, const TiffDecode& f,
                               cudaStream_t stream);
cudaError_t launch_synth(uint32_t* grid, const SynthSpec& spec, cudaStream_t stream);
cudaError_t launch_count_black(const uint32_t* grid, uint64_t n, uint64_t stride, unsigned long long* out,
                               cudaStream_t stream);
// DDA skip table over 2^shift-voxel bricks with radii up to `cap` (layout: xn_device.cuh SkipTable);
// *table_out is cudaMalloc'ed, dims_out = table extent in bricks including the one-brick border,
// *uniform_fraction_out (nullable) = share of the grid's bricks that hold one colour
cudaError_t build_skip_table(const uint32_t* grid, uint32_t nx, uint32_t ny, uint32_t nz, uint32_t shift, uint32_t cap,
                             uint4** table_out, uint32_t dims_out[3], double* uniform_fraction_out, cudaStream_t stream);
// out2[0] += bits set, out2[1] += non-zero bytes of a touch map of `words` 32-bit words
cudaError_t launch_touch_count(const uint32_t* bits, uint64_t words, unsigned long long* out2, cudaStream_t stream);
cudaError_t launch_stats_totals(const uint32_t* steps, const unsigned long long* bytes, uint64_t n,
                                unsigned long long* totals, cudaStream_t stream);
} // namespace xn
