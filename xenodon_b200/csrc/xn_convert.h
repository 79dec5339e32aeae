// xn_convert.h -- GPU octree construction (xn_convert.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "xenodon_b200.h"

namespace xn {
// Builds the octree of an RGBA8 grid resident on the current device.  *d_nodes_out receives a
// cudaMalloc'ed array of *count_out 40-byte nodes (the .svo node array, node 0 = root).
// heuristic 0 = --chan-diff (param 0..255), 1 = --std-dev (param >= 0; XN_ERR_LIMIT when the threshold
// is within rounding distance of a cell's deviation: the host builder decides those).
void gpu_build_octree(const uint32_t* d_grid, uint64_t nx, uint64_t ny, uint64_t nz, int heuristic, double param, bool rope,
                      cudaStream_t stream, void** d_nodes_out, uint64_t* count_out, uint64_t* side_out,
                      xn_build_stats* stats_out);
// `--dag`: merges identical subtrees of a sparse (non-rope) tree built by gpu_build_octree, byte-identical
// to the reference's HashCache builder (src/model/OctreeConstruction.h:19-30).  d_sparse stays owned
// by the caller; *d_dag_out is cudaMalloc'ed.
void gpu_dag_from_sparse(const void* d_sparse, uint64_t count, cudaStream_t stream, void** d_dag_out, uint64_t* dag_count_out,
                         uint64_t* unique_leaves_out);
} // namespace xn
