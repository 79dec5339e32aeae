// xn_dag.cu -- `xenodon convert --dag` on the GPU: merges identical subtrees of a sparse octree
// into a directed acyclic graph, byte-identical to the reference's HashCache builder
// (reference src/model/OctreeConstruction.h:19-30, :66-112).
//
// The reference visits the FULL tree in post-order whatever the cache says (construct() recurses
// before it inserts, :170-189); a node equal to one inserted earlier -- same eight child indices,
// colour, leaf flag and depth -- gets that node's index instead of a new one, and the array is
// reversed at the end.  In the sparse array this builder starts from (the reversed post-order),
// "inserted earlier" means "larger index".  So the DAG is exactly the sparse array FILTERED to one
// representative per class of identical subtrees -- the member with the largest sparse index --
// in unchanged order, with child pointers redirected to the representatives' new positions.
//
// Classes are found bottom-up, one tree level at a time (children are one level deeper, so their
// classes are known): a node's key is (its children's representatives, colour, is_leaf_depth); nodes
// of a level are sorted by a 64-bit hash of the key with a stable radix sort over a list kept in
// descending index order, so the head of every run of equal hashes is the representative; keys
// are then compared word for word along each run (a hash collision is reported, never merged).
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <string>
#include <vector>

#include "host/xn_host.hpp"
#include "xn_convert.h"

namespace xn {
namespace {

constexpr uint32_t LEAF_BIT = 0x80000000u;

void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(XN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

struct Buffers {
    std::vector<void*> ptrs;
    ~Buffers() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    T* alloc(uint64_t n) {
        void* p = nullptr;
        check(cudaMalloc(&p, std::max<uint64_t>(n, 1) * sizeof(T)), "cudaMalloc (dag)");
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
};

unsigned grid_for(uint64_t n) { return (unsigned)std::min<uint64_t>(std::max<uint64_t>((n + 255) / 256, 1), 148ull * 32); }

__device__ __forceinline__ uint64_t mix64(uint64_t h, uint64_t v) { // splitmix-style combine
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= h >> 30;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 27;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 31;
    return h;
}

// list of node indices in DESCENDING order with the node's depth as sort key
__global__ void dag_list_kernel(const uint32_t* __restrict__ nodes, uint64_t count, uint8_t* __restrict__ keys,
                                uint32_t* __restrict__ vals, uint32_t* __restrict__ level_count) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t f = (uint32_t)(count - 1 - j);
        const uint32_t depth = min(nodes[(uint64_t)f * 10u + 9u] & 0x7FFFFFFFu, 31u);
        keys[j] = (uint8_t)depth;
        vals[j] = f;
        atomicAdd(&level_count[depth], 1u);
    }
}

// the key of node f: children's representatives (0 for a leaf), colour, is_leaf_depth
__device__ __forceinline__ void dag_key(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ rep, uint32_t f,
                                        uint32_t key[10]) {
    const uint32_t* n = nodes + (uint64_t)f * 10u;
    const bool leaf = (n[9] & LEAF_BIT) != 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) key[k] = leaf ? n[k] : rep[n[k]];
    key[8] = n[8];
    key[9] = n[9];
}

__global__ void dag_hash_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ rep,
                                const uint32_t* __restrict__ items, uint64_t m, uint64_t* __restrict__ hashes) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t key[10];
        dag_key(nodes, rep, items[j], key);
        uint64_t h = 0x243F6A8885A308D3ull;
#pragma unroll
        for (int k = 0; k < 10; k += 2) h = mix64(h, (uint64_t)key[k] | ((uint64_t)key[k + 1] << 32));
        hashes[j] = h;
    }
}

// head flag of every run of equal hashes (as the run's own position, 0 elsewhere, for a max-scan);
// inside a run every key must equal its predecessor's -- otherwise two different keys share a hash
__global__ void dag_heads_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ rep,
                                 const uint64_t* __restrict__ hashes, const uint32_t* __restrict__ items, uint64_t m,
                                 uint32_t* __restrict__ head_pos, unsigned long long* __restrict__ collisions) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x) {
        const bool head = j == 0 || hashes[j] != hashes[j - 1];
        head_pos[j] = head ? (uint32_t)j : 0u;
        if (!head) {
            uint32_t a[10], b[10];
            dag_key(nodes, rep, items[j], a);
            dag_key(nodes, rep, items[j - 1], b);
            bool same = true;
#pragma unroll
            for (int k = 0; k < 10; ++k) same &= a[k] == b[k];
            if (!same) atomicAdd(collisions, 1ull);
        }
    }
}

__global__ void dag_assign_kernel(const uint32_t* __restrict__ items, const uint32_t* __restrict__ head_pos, uint64_t m,
                                  uint32_t* __restrict__ rep) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (uint64_t)gridDim.x * blockDim.x)
        rep[items[j]] = items[head_pos[j]]; // the run's head has the largest index: inserted first by the reference
}

__global__ void dag_flag_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ rep, uint64_t count,
                                uint32_t* __restrict__ is_rep, unsigned long long* __restrict__ unique_leaves) {
    unsigned long long leaves = 0;
    for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < count; f += (uint64_t)gridDim.x * blockDim.x) {
        const bool r = rep[f] == (uint32_t)f;
        is_rep[f] = r ? 1u : 0u;
        leaves += r && (nodes[f * 10u + 9u] & LEAF_BIT) != 0u;
    }
    for (int o = 16; o > 0; o >>= 1) leaves += __shfl_xor_sync(0xFFFFFFFFu, leaves, o);
    if ((threadIdx.x & 31) == 0 && leaves) atomicAdd(unique_leaves, leaves);
}

__global__ void dag_emit_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ rep,
                                const uint32_t* __restrict__ is_rep, const uint32_t* __restrict__ new_index, uint64_t count,
                                uint32_t* __restrict__ out) {
    for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < count; f += (uint64_t)gridDim.x * blockDim.x) {
        if (!is_rep[f]) continue;
        const uint32_t* n = nodes + f * 10u;
        uint32_t* o = out + (uint64_t)new_index[f] * 10u;
        const bool leaf = (n[9] & LEAF_BIT) != 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = leaf ? n[k] : new_index[rep[n[k]]];
        o[8] = n[8];
        o[9] = n[9];
    }
}

} // namespace

void gpu_dag_from_sparse(const void* d_sparse, uint64_t count, cudaStream_t stream, void** d_dag_out, uint64_t* dag_count_out,
                         uint64_t* unique_leaves_out) {
    if (!d_sparse || count == 0) throw Error(XN_ERR_INVALID, "gpu_dag_from_sparse: empty tree");
    if (count > 0x7FFFFFFFull) throw Error(XN_ERR_LIMIT, "gpu_dag_from_sparse: tree too large");
    const uint32_t* nodes = static_cast<const uint32_t*>(d_sparse);
    Buffers buf;
    uint8_t* k_in = buf.alloc<uint8_t>(count);
    uint8_t* k_out = buf.alloc<uint8_t>(count);
    uint32_t* v_in = buf.alloc<uint32_t>(count);
    uint32_t* list = buf.alloc<uint32_t>(count); // node indices by level, descending inside a level
    uint32_t* rep = buf.alloc<uint32_t>(count);
    uint32_t* d_levels = buf.alloc<uint32_t>(32);
    unsigned long long* d_counters = buf.alloc<unsigned long long>(2); // [0] collisions, [1] unique leaves
    check(cudaMemsetAsync(d_levels, 0, 32 * 4, stream), "cudaMemsetAsync");
    check(cudaMemsetAsync(d_counters, 0, 16, stream), "cudaMemsetAsync");
    dag_list_kernel<<<grid_for(count), 256, 0, stream>>>(nodes, count, k_in, v_in, d_levels);
    check(cudaGetLastError(), "dag_list_kernel");
    size_t tmp_bytes = 0;
    check(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, list, (int64_t)count, 0, 5, stream), "cub size");
    uint32_t levels[32];
    check(cudaMemcpyAsync(levels, d_levels, sizeof levels, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    uint64_t largest = 0;
    for (uint32_t c : levels) largest = std::max<uint64_t>(largest, c);
    // per-level scratch, sized for the largest level
    uint64_t* h_in = buf.alloc<uint64_t>(largest);
    uint64_t* h_out = buf.alloc<uint64_t>(largest);
    uint32_t* items = buf.alloc<uint32_t>(largest);
    uint32_t* head_pos = buf.alloc<uint32_t>(largest);
    uint32_t* head_scan = buf.alloc<uint32_t>(largest);
    size_t tmp2 = 0, tmp3 = 0;
    check(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, h_in, h_out, list, items, (int64_t)largest, 0, 64, stream), "cub size");
    check(cub::DeviceScan::InclusiveScan(nullptr, tmp3, head_pos, head_scan, cub::Max(), (int64_t)largest, stream), "cub size");
    uint32_t* is_rep = buf.alloc<uint32_t>(count);
    uint32_t* new_index = buf.alloc<uint32_t>(count + 1);
    size_t tmp4 = 0;
    check(cub::DeviceScan::ExclusiveSum(nullptr, tmp4, is_rep, new_index, (int64_t)count, stream), "cub size");
    const size_t tmp_all = std::max({tmp_bytes, tmp2, tmp3, tmp4});
    void* tmp = buf.alloc<uint8_t>(tmp_all);

    size_t t = tmp_all;
    check(cub::DeviceRadixSort::SortPairs(tmp, t, k_in, k_out, v_in, list, (int64_t)count, 0, 5, stream), "cub sort (levels)");

    // bottom-up over the levels; level d occupies list[offset[d] .. offset[d] + levels[d])
    uint64_t offset[33];
    offset[0] = 0;
    for (int dpt = 0; dpt < 32; ++dpt) offset[dpt + 1] = offset[dpt] + levels[dpt];
    for (int dpt = 31; dpt >= 0; --dpt) {
        const uint64_t m = levels[dpt];
        if (m == 0) continue;
        const uint32_t* level_items = list + offset[dpt];
        dag_hash_kernel<<<grid_for(m), 256, 0, stream>>>(nodes, rep, level_items, m, h_in);
        t = tmp_all;
        check(cub::DeviceRadixSort::SortPairs(tmp, t, h_in, h_out, level_items, items, (int64_t)m, 0, 64, stream), "cub sort (hashes)");
        dag_heads_kernel<<<grid_for(m), 256, 0, stream>>>(nodes, rep, h_out, items, m, head_pos, d_counters);
        t = tmp_all;
        check(cub::DeviceScan::InclusiveScan(tmp, t, head_pos, head_scan, cub::Max(), (int64_t)m, stream), "cub scan (heads)");
        dag_assign_kernel<<<grid_for(m), 256, 0, stream>>>(items, head_scan, m, rep);
        check(cudaGetLastError(), "dag level kernels");
    }
    dag_flag_kernel<<<grid_for(count), 256, 0, stream>>>(nodes, rep, count, is_rep, d_counters + 1);
    t = tmp_all;
    check(cub::DeviceScan::ExclusiveSum(tmp, t, is_rep, new_index, (int64_t)count, stream), "cub scan (indices)");
    unsigned long long counters[2] = {0, 0};
    uint32_t last_index = 0, last_flag = 0;
    check(cudaMemcpyAsync(counters, d_counters, 16, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    check(cudaMemcpyAsync(&last_index, new_index + (count - 1), 4, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    check(cudaMemcpyAsync(&last_flag, is_rep + (count - 1), 4, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    if (counters[0] != 0)
        throw Error(XN_ERR_LIMIT, "gpu_dag_from_sparse: two different subtrees share a 64-bit hash (use the host builder)");
    const uint64_t unique = (uint64_t)last_index + last_flag;
    uint32_t* d_dag = nullptr;
    check(cudaMalloc((void**)&d_dag, unique * 40), "cudaMalloc (dag nodes)");
    dag_emit_kernel<<<grid_for(count), 256, 0, stream>>>(nodes, rep, is_rep, new_index, count, d_dag);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
        cudaFree(d_dag);
        check(e, "dag_emit_kernel");
    }
    *d_dag_out = d_dag;
    *dag_count_out = unique;
    if (unique_leaves_out) *unique_leaves_out = counters[1];
}

} // namespace xn
