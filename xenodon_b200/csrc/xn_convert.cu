// xn_convert.cu -- `xenodon convert` on the GPU: builds the sparse
// voxel octree of a grid resident in HBM, byte-identical to the reference's recursive builder
// (reference src/model/OctreeConstruction.h:124-237, Grid::vol_scan src/model/Grid.cpp:81-137,
// Octree::generate_ropes src/model/Octree.cpp:181-201) for the --chan-diff heuristic and the
// sparse / rope tree types.
//
// The reference rescans every region at every level (O(N * depth) voxel reads).  Here:
//   1. bottom-up: one pass per level builds a min / max / sum pyramid and, from it, every
//      cell's verdict (leaf or interior; vol_scan's max_diff > threshold) and subtree size;
//   2. top-down: one pass per level turns subtree sizes into node indices.  The reference
//      inserts nodes post-order and reverses the array (OctreeConstruction.h:92-112), which is
//      a pre-order walk visiting children 7..0, so
//          index(child k) = index(parent) + 1 + sum_{j > k} size(child j),
//      and writes the 40-byte node records (children mirrored into the reversed numbering);
//   3. (--rope) one pass over the leaves runs the six Octree::find descents per leaf.
// All arithmetic is integer; every voxel is read once.  --std-dev is decided from exact integer
// sums with a rigorous bound on the reference's binary64 rounding (SplitRule below); DAG merging
// (--dag) is a pass over the finished sparse array (xn_dag.cu).
#include <algorithm>
#include <vector>

#include "host/xn_host.hpp"
#include "xn_convert.h"

namespace xn {
namespace {

constexpr uint32_t TOP = 0x80000000u;   // size/index word: top bit set = "not an existing interior node"
constexpr uint32_t LEAF_BIT = 0x80000000u;

struct Dims {
    uint32_t nx, ny, nz;
};

__device__ __forceinline__ uint32_t min4(uint32_t a, uint32_t b) { return __vminu4(a, b); }
__device__ __forceinline__ uint32_t max4(uint32_t a, uint32_t b) { return __vmaxu4(a, b); }

// clipped voxel count of the cell [off, off + e)^3
__device__ __forceinline__ uint64_t clipped_count(Dims d, uint32_t ox, uint32_t oy, uint32_t oz, uint32_t e) {
    const uint64_t x = min(d.nx, ox + e) - min(d.nx, ox);
    const uint64_t y = min(d.ny, oy + e) - min(d.ny, oy);
    const uint64_t z = min(d.nz, oz + e) - min(d.nz, oz);
    return x * y * z;
}

template <typename SumT>
__device__ __forceinline__ uint32_t avg_color(const SumT s[4], uint64_t n) {
    if (n == 0) return 0u;
    return (uint32_t)(s[0] / n) | ((uint32_t)(s[1] / n) << 8) | ((uint32_t)(s[2] / n) << 16) |
           ((uint32_t)(s[3] / n) << 24);
}

__device__ __forceinline__ uint32_t max_diff(uint32_t mn, uint32_t mx) {
    const uint32_t d = __vsubus4(mx, mn);
    return max(max(d & 0xFFu, (d >> 8) & 0xFFu), max((d >> 16) & 0xFFu, d >> 24));
}

// One pyramid level: per cell min / max (packed RGBA), per-channel sums and the size/index word.
template <typename SumT>
struct Level {
    uint32_t* mn;
    uint32_t* mx;
    SumT* sum; // 4 per cell
    uint32_t* si;
    uint32_t cells; // per axis
    unsigned long long* sq; // --std-dev only: sum over the cell's voxels and channels of x^2 (else nullptr)
};

// Split rule of the two heuristics (src/model/OctreeConstruction.h:32-48).
//   --chan-diff: vol_scan's max channel difference > threshold (integers).
//   --std-dev  : stddev_scan's sqrt(sum((x - avg)^2) / n) > threshold, which the reference evaluates
//     in binary64 with one rounded addition per voxel (src/model/Grid.cpp:139-214).  The exact
//     value follows from integer sums, V = (n Q - sum_c A_c^2) / n^2 with A_c = sum x_c and
//     Q = sum x^2; the reference's rounded result differs from sqrt(V) by a relative error of at
//     most tau = (n + 16) 2^-54 (n - 1 additions of non-negative terms, a handful of roundings per
//     term, one division, one square root).  The verdict is therefore CERTAIN unless V lies within
//     a factor 1 +- 8 tau of threshold^2; such a cell is counted in *uncertain and the caller gives
//     the volume to the host builder, which replays the reference's additions.  threshold = 0 is
//     exact: the deviation is 0 iff every voxel equals the average.
struct SplitRule {
    int heuristic;      // 0 --chan-diff, 1 --std-dev
    uint32_t chan_diff;
    double std_dev;
};
template <typename SumT>
__device__ __forceinline__ bool wants_split(const SplitRule& rule, uint32_t mn, uint32_t mx, const SumT s[4],
                                            unsigned long long q, uint64_t n, unsigned long long* uncertain) {
    if (rule.heuristic == 0) return max_diff(mn, mx) > rule.chan_diff;
    if (n == 0) return false;
    unsigned __int128 t = (unsigned __int128)n * q;
    for (int c = 0; c < 4; ++c) t -= (unsigned __int128)(uint64_t)s[c] * (uint64_t)s[c]; // n Q >= sum A_c^2 (Cauchy-Schwarz)
    if (rule.std_dev == 0.0) return t != 0;
    const double td = (double)(uint64_t)(t >> 64) * 18446744073709551616.0 + (double)(uint64_t)t;
    const double nd = (double)n;
    const double v = td / (nd * nd);
    const double thr2 = rule.std_dev * rule.std_dev;
    const double tau = (nd + 16.0) * 5.551115123125783e-17; // 2^-54
    if (v > thr2 * (1.0 + 8.0 * tau)) return true;
    if (v < thr2 * (1.0 - 8.0 * tau)) return false;
    atomicAdd(uncertain, 1ull);
    return false;
}

// ---- bottom-up, level 1: children are voxels ----
__global__ void pyramid_level1(const uint32_t* __restrict__ grid, Dims d, Level<uint32_t> out, SplitRule rule,
                               unsigned long long* uncertain) {
    const uint64_t total = (uint64_t)out.cells * out.cells * out.cells;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t cx = (uint32_t)(c % out.cells), cy = (uint32_t)((c / out.cells) % out.cells),
                       cz = (uint32_t)(c / ((uint64_t)out.cells * out.cells));
        const uint32_t ox = cx * 2, oy = cy * 2, oz = cz * 2;
        uint32_t mn = 0xFFFFFFFFu, mx = 0u, s[4] = {0, 0, 0, 0};
        unsigned long long q = 0;
        if (ox < d.nx && oy < d.ny && oz < d.nz) {
            for (uint32_t z = oz; z < min(oz + 2, d.nz); ++z)
                for (uint32_t y = oy; y < min(oy + 2, d.ny); ++y)
                    for (uint32_t x = ox; x < min(ox + 2, d.nx); ++x) {
                        const uint32_t v = grid[(uint64_t)x + (uint64_t)y * d.nx + (uint64_t)z * d.nx * d.ny];
                        mn = min4(mn, v);
                        mx = max4(mx, v);
                        s[0] += v & 0xFFu;
                        s[1] += (v >> 8) & 0xFFu;
                        s[2] += (v >> 16) & 0xFFu;
                        s[3] += v >> 24;
                        q += (v & 0xFFu) * (v & 0xFFu) + ((v >> 8) & 0xFFu) * ((v >> 8) & 0xFFu) +
                             ((v >> 16) & 0xFFu) * ((v >> 16) & 0xFFu) + (v >> 24) * (v >> 24);
                    }
            const bool fully_in = ox + 2 <= d.nx && oy + 2 <= d.ny && oz + 2 <= d.nz;
            const bool leaf = !wants_split(rule, mn, mx, s, q, clipped_count(d, ox, oy, oz, 2), uncertain) && fully_in;
            out.si[c] = TOP | (leaf ? 1u : 9u);
        } else {
            out.si[c] = TOP | 1u; // outside the source grid: black leaf
        }
        out.mn[c] = mn;
        out.mx[c] = mx;
        out.sum[4 * c + 0] = s[0];
        out.sum[4 * c + 1] = s[1];
        out.sum[4 * c + 2] = s[2];
        out.sum[4 * c + 3] = s[3];
        if (out.sq) out.sq[c] = q;
    }
}

// ---- bottom-up, level L >= 2: children are cells of level L-1 ----
template <typename ChildSumT, typename SumT>
__global__ void pyramid_level(Level<ChildSumT> in, Level<SumT> out, Dims d, uint32_t extent, SplitRule rule,
                              unsigned long long* overflow, unsigned long long* uncertain) {
    const uint64_t total = (uint64_t)out.cells * out.cells * out.cells;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t cx = (uint32_t)(c % out.cells), cy = (uint32_t)((c / out.cells) % out.cells),
                       cz = (uint32_t)(c / ((uint64_t)out.cells * out.cells));
        const uint32_t ox = cx * extent, oy = cy * extent, oz = cz * extent;
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        SumT s[4] = {0, 0, 0, 0};
        unsigned long long q = 0;
        uint32_t word = TOP | 1u;
        if (ox < d.nx && oy < d.ny && oz < d.nz) {
            uint64_t size = 1;
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t kx = 2 * cx + ((k >> 2) & 1u), ky = 2 * cy + ((k >> 1) & 1u), kz = 2 * cz + (k & 1u);
                const uint64_t kc = (uint64_t)kx + (uint64_t)ky * in.cells + (uint64_t)kz * in.cells * in.cells;
                size += in.si[kc] & ~TOP;
                mn = min4(mn, in.mn[kc]); // cells outside the grid hold the neutral elements
                mx = max4(mx, in.mx[kc]);
                for (int ch = 0; ch < 4; ++ch) s[ch] += (SumT)in.sum[4 * kc + ch];
                if (in.sq) q += in.sq[kc];
            }
            const bool fully_in = ox + extent <= d.nx && oy + extent <= d.ny && oz + extent <= d.nz;
            const bool leaf =
                !wants_split(rule, mn, mx, s, q, clipped_count(d, ox, oy, oz, extent), uncertain) && fully_in;
            if (leaf) size = 1;
            if (size >= TOP) atomicAdd(overflow, 1ull); // more than 2^31 - 1 nodes: not representable here
            word = TOP | (uint32_t)size;
        }
        out.si[c] = word;
        out.mn[c] = mn;
        out.mx[c] = mx;
        for (int ch = 0; ch < 4; ++ch) out.sum[4 * c + ch] = s[ch];
        if (out.sq) out.sq[c] = q;
    }
}

__device__ __forceinline__ void write_node(uint32_t* __restrict__ nodes, uint32_t index, const uint32_t ch[8],
                                           uint32_t color, uint32_t is_leaf_depth) {
    uint32_t* n = nodes + (uint64_t)index * 10u;
    // 40-byte records are 8-byte aligned: five 64-bit stores
    uint2* n2 = reinterpret_cast<uint2*>(n);
    n2[0] = make_uint2(ch[0], ch[1]);
    n2[1] = make_uint2(ch[2], ch[3]);
    n2[2] = make_uint2(ch[4], ch[5]);
    n2[3] = make_uint2(ch[6], ch[7]);
    n2[4] = make_uint2(color, is_leaf_depth);
}

// ---- top-down: parents at level L (extent e), children at level L-1 (cells) ----
template <typename ChildSumT, typename SumT>
__global__ void emit_level(Level<SumT> par, Level<ChildSumT> chl, Dims d, uint32_t extent, uint32_t depth,
                           uint32_t* __restrict__ nodes, uint32_t* __restrict__ level_used) {
    const uint64_t total = (uint64_t)par.cells * par.cells * par.cells;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t index = par.si[c];
        if (index & TOP) continue; // not an existing interior node
        const uint32_t cx = (uint32_t)(c % par.cells), cy = (uint32_t)((c / par.cells) % par.cells),
                       cz = (uint32_t)(c / ((uint64_t)par.cells * par.cells));
        const uint32_t half = extent / 2;
        uint32_t cidx[8];
        uint64_t kcell[8];
        uint32_t ksize[8];
        for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t kx = 2 * cx + ((k >> 2) & 1u), ky = 2 * cy + ((k >> 1) & 1u), kz = 2 * cz + (k & 1u);
            kcell[k] = (uint64_t)kx + (uint64_t)ky * chl.cells + (uint64_t)kz * chl.cells * chl.cells;
            ksize[k] = chl.si[kcell[k]] & ~TOP;
        }
        uint32_t running = index + 1u;
        for (int k = 7; k >= 0; --k) {
            cidx[k] = running;
            running += ksize[k];
        }
        SumT ps[4] = {par.sum[4 * c], par.sum[4 * c + 1], par.sum[4 * c + 2], par.sum[4 * c + 3]};
        write_node(nodes, index, cidx, avg_color(ps, clipped_count(d, cx * extent, cy * extent, cz * extent, extent)),
                   depth);
        const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t k = 0; k < 8; ++k) {
            if (ksize[k] == 1u) { // leaf child (possibly outside the grid: black, n = 0)
                const uint32_t kx = 2 * cx + ((k >> 2) & 1u), ky = 2 * cy + ((k >> 1) & 1u), kz = 2 * cz + (k & 1u);
                ChildSumT cs[4] = {chl.sum[4 * kcell[k]], chl.sum[4 * kcell[k] + 1], chl.sum[4 * kcell[k] + 2],
                                   chl.sum[4 * kcell[k] + 3]};
                const uint64_t n = clipped_count(d, kx * half, ky * half, kz * half, half);
                write_node(nodes, cidx[k], zero, avg_color(cs, n), LEAF_BIT | (depth + 1u));
                chl.si[kcell[k]] = 0xFFFFFFFFu;
            } else {
                chl.si[kcell[k]] = cidx[k];
            }
        }
        *level_used = 1u; // depth + 1 is populated
    }
}

// ---- top-down, last step: parents at level 1, children are voxels ----
__global__ void emit_voxels(Level<uint32_t> par, const uint32_t* __restrict__ grid, Dims d, uint32_t depth,
                            uint32_t* __restrict__ nodes, uint32_t* __restrict__ level_used) {
    const uint64_t total = (uint64_t)par.cells * par.cells * par.cells;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t index = par.si[c];
        if (index & TOP) continue;
        const uint32_t cx = (uint32_t)(c % par.cells), cy = (uint32_t)((c / par.cells) % par.cells),
                       cz = (uint32_t)(c / ((uint64_t)par.cells * par.cells));
        uint32_t cidx[8];
        for (int k = 7, running = (int)index + 1; k >= 0; --k) cidx[k] = (uint32_t)running++;
        uint32_t ps[4] = {par.sum[4 * c], par.sum[4 * c + 1], par.sum[4 * c + 2], par.sum[4 * c + 3]};
        write_node(nodes, index, cidx, avg_color(ps, clipped_count(d, cx * 2, cy * 2, cz * 2, 2)), depth);
        const uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t x = 2 * cx + ((k >> 2) & 1u), y = 2 * cy + ((k >> 1) & 1u), z = 2 * cz + (k & 1u);
            uint32_t color = 0;
            if (x < d.nx && y < d.ny && z < d.nz) color = grid[(uint64_t)x + (uint64_t)y * d.nx + (uint64_t)z * d.nx * d.ny];
            write_node(nodes, cidx[k], zero, color, LEAF_BIT | (depth + 1u));
        }
        *level_used = 1u;
    }
}

// ---- ropes: Octree::find (src/model/Octree.cpp:116-153) on the emitted array ----
__device__ uint32_t find_node(const uint32_t* __restrict__ nodes, uint64_t dim, uint64_t px, uint64_t py, uint64_t pz,
                              uint32_t max_depth) {
    uint64_t extent = dim;
    if (px >= extent || py >= extent || pz >= extent) return 0u; // also catches the wrapped "-extent"
    uint32_t index = 0;
    uint64_t ox = 0, oy = 0, oz = 0;
    for (;;) {
        extent /= 2;
        const uint32_t* n = nodes + (uint64_t)index * 10u;
        if ((n[9] & LEAF_BIT) || extent == 0 || max_depth == 0) return index;
        uint32_t ci = 0;
        if (px >= ox + extent) { ci |= 4u; ox += extent; }
        if (py >= oy + extent) { ci |= 2u; oy += extent; }
        if (pz >= oz + extent) { ci |= 1u; oz += extent; }
        index = n[ci];
        --max_depth;
    }
}

__device__ __forceinline__ void write_ropes(uint32_t* __restrict__ nodes, uint32_t leaf, uint64_t dim, uint64_t x,
                                            uint64_t y, uint64_t z, uint64_t e, uint32_t depth) {
    uint32_t* n = nodes + (uint64_t)leaf * 10u;
    n[0] = find_node(nodes, dim, x + e, y, z, depth);
    n[1] = find_node(nodes, dim, x - e, y, z, depth); // unsigned wrap = out of range, as in the reference
    n[2] = find_node(nodes, dim, x, y + e, z, depth);
    n[3] = find_node(nodes, dim, x, y - e, z, depth);
    n[4] = find_node(nodes, dim, x, y, z + e, depth);
    n[5] = find_node(nodes, dim, x, y, z - e, depth);
}

// leaves are visited through their parents: every existing interior cell of level L handles
// those of its children that are leaves (child extent = extent / 2, depth + 1)
__global__ void rope_level(const uint32_t* __restrict__ par_si, uint32_t cells, uint32_t extent, uint32_t depth,
                           uint64_t dim, uint32_t* __restrict__ nodes) {
    const uint64_t total = (uint64_t)cells * cells * cells;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t index = par_si[c];
        if (index & TOP) continue;
        const uint32_t cx = (uint32_t)(c % cells), cy = (uint32_t)((c / cells) % cells),
                       cz = (uint32_t)(c / ((uint64_t)cells * cells));
        const uint32_t half = extent / 2;
        const uint32_t* pn = nodes + (uint64_t)index * 10u;
        for (uint32_t k = 0; k < 8; ++k) {
            const uint32_t child = pn[k];
            if (!(nodes[(uint64_t)child * 10u + 9u] & LEAF_BIT)) continue;
            const uint64_t x = (uint64_t)cx * extent + ((k >> 2) & 1u) * half;
            const uint64_t y = (uint64_t)cy * extent + ((k >> 1) & 1u) * half;
            const uint64_t z = (uint64_t)cz * extent + (k & 1u) * half;
            write_ropes(nodes, child, dim, x, y, z, half, depth + 1u);
        }
    }
}

uint64_t ceil_2pow(uint64_t x) {
    --x;
    x |= x >> 1;
    x |= x >> 2;
    x |= x >> 4;
    x |= x >> 8;
    x |= x >> 16;
    x |= x >> 32;
    return ++x;
}

struct DeviceBuffers {
    std::vector<void*> ptrs;
    ~DeviceBuffers() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <typename T>
    T* alloc(uint64_t n) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<uint64_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) throw Error(XN_ERR_CUDA, std::string("cudaMalloc (octree pyramid): ") + cudaGetErrorString(e));
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
};

void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(XN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int blocks_for(uint64_t cells_total) {
    const uint64_t b = (cells_total + 255) / 256;
    return (int)std::min<uint64_t>(std::max<uint64_t>(b, 1), 148ull * 32);
}

} // namespace

void gpu_build_octree(const uint32_t* d_grid, uint64_t nx, uint64_t ny, uint64_t nz, int heuristic, double param, bool rope,
                      cudaStream_t stream, void** d_nodes_out, uint64_t* count_out, uint64_t* side_out,
                      xn_build_stats* stats_out) {
    if (heuristic != 0 && heuristic != 1) throw Error(XN_ERR_INVALID, "gpu_build_octree: unknown heuristic");
    if (heuristic == 0 && !(param >= 0.0 && param <= 255.0)) throw Error(XN_ERR_INVALID, "channel difference must be 0..255");
    if (heuristic == 1 && !(param >= 0.0)) throw Error(XN_ERR_INVALID, "standard deviation must be >= 0");
    const SplitRule rule{heuristic, heuristic == 0 ? (uint32_t)param : 0u, heuristic == 1 ? param : 0.0};
    const bool want_sq = heuristic == 1;
    if (!d_grid || nx == 0 || ny == 0 || nz == 0) throw Error(XN_ERR_INVALID, "gpu_build_octree: empty grid");
    const uint64_t dim = std::max({ceil_2pow(nx), ceil_2pow(ny), ceil_2pow(nz)});
    if (dim > 65536) throw Error(XN_ERR_LIMIT, "gpu_build_octree: grid too large");
    const Dims d{(uint32_t)nx, (uint32_t)ny, (uint32_t)nz};
    int lmax = 0;
    while ((1ull << lmax) < dim) ++lmax; // level lmax = root (extent dim, depth 0)
    *side_out = dim;

    xn_build_stats stats{0, 0, 0, 0};
    if (lmax == 0) { // a 1x1x1 grid: the root is the voxel
        uint32_t v = 0;
        check(cudaMemcpyAsync(&v, d_grid, 4, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
        check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
        const uint32_t rec[10] = {0, 0, 0, 0, 0, 0, 0, 0, v, LEAF_BIT};
        void* dn = nullptr;
        check(cudaMalloc(&dn, 40), "cudaMalloc");
        check(cudaMemcpyAsync(dn, rec, 40, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync");
        check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
        *d_nodes_out = dn;
        *count_out = 1;
        stats.total_leaves = stats.unique_leaves = stats.total_nodes = 1;
        if (stats_out) *stats_out = stats;
        return;
    }

    // per-channel sums fit 32 bits up to level 8 (2^24 voxels x 255); above that the handful of
    // cells use 64-bit sums
    constexpr int WIDE_FROM = 9;
    DeviceBuffers buf;
    std::vector<Level<uint32_t>> lo(lmax + 1);
    std::vector<Level<uint64_t>> hi(lmax + 1);
    for (int L = 1; L <= lmax; ++L) {
        const uint32_t cells = (uint32_t)(dim >> L);
        const uint64_t total = (uint64_t)cells * cells * cells;
        unsigned long long* sq = want_sq ? buf.alloc<unsigned long long>(total) : nullptr;
        if (L < WIDE_FROM)
            lo[L] = Level<uint32_t>{buf.alloc<uint32_t>(total), buf.alloc<uint32_t>(total), buf.alloc<uint32_t>(4 * total),
                                    buf.alloc<uint32_t>(total), cells, sq};
        else
            hi[L] = Level<uint64_t>{buf.alloc<uint32_t>(total), buf.alloc<uint32_t>(total), buf.alloc<uint64_t>(4 * total),
                                    buf.alloc<uint32_t>(total), cells, sq};
    }
    unsigned long long* d_overflow = buf.alloc<unsigned long long>(2); // [0] node-count overflow, [1] uncertain verdicts
    unsigned long long* d_uncertain = d_overflow + 1;
    uint32_t* d_used = buf.alloc<uint32_t>(lmax + 2);
    check(cudaMemsetAsync(d_overflow, 0, 16, stream), "cudaMemsetAsync");
    check(cudaMemsetAsync(d_used, 0, (lmax + 2) * 4, stream), "cudaMemsetAsync");

    // 1. bottom-up
    pyramid_level1<<<blocks_for((uint64_t)lo[1].cells * lo[1].cells * lo[1].cells), 256, 0, stream>>>(d_grid, d, lo[1], rule,
                                                                                                 d_uncertain);
    for (int L = 2; L <= lmax; ++L) {
        const uint32_t extent = 1u << L;
        const uint32_t cells = (uint32_t)(dim >> L);
        const int blocks = blocks_for((uint64_t)cells * cells * cells);
        if (L < WIDE_FROM)
            pyramid_level<uint32_t, uint32_t><<<blocks, 256, 0, stream>>>(lo[L - 1], lo[L], d, extent, rule, d_overflow, d_uncertain);
        else if (L == WIDE_FROM)
            pyramid_level<uint32_t, uint64_t><<<blocks, 256, 0, stream>>>(lo[L - 1], hi[L], d, extent, rule, d_overflow, d_uncertain);
        else
            pyramid_level<uint64_t, uint64_t><<<blocks, 256, 0, stream>>>(hi[L - 1], hi[L], d, extent, rule, d_overflow, d_uncertain);
    }
    check(cudaGetLastError(), "pyramid kernels");

    // root verdict
    uint32_t root_word = 0, root_mn = 0;
    unsigned long long overflow[2] = {0, 0};
    uint32_t* root_si = lmax < WIDE_FROM ? lo[lmax].si : hi[lmax].si;
    check(cudaMemcpyAsync(&root_word, root_si, 4, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    check(cudaMemcpyAsync(overflow, d_overflow, 16, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    (void)root_mn;
    if (overflow[1])
        throw Error(XN_ERR_LIMIT, "--std-dev: the threshold lies within rounding distance of " + std::to_string(overflow[1]) +
                                      " cell deviation(s); only the host builder replays the reference's additions");
    if (overflow[0]) throw Error(XN_ERR_LIMIT, "octree exceeds 2^31 - 1 nodes (GPU builder limit)");
    const uint64_t count = root_word & ~TOP;

    uint32_t* d_nodes = nullptr;
    check(cudaMalloc((void**)&d_nodes, count * 40), "cudaMalloc (octree nodes)");
    try {
        if (count == 1) {
            // the whole cube is one leaf: colour = average over the (full) grid
            uint64_t sums64[4] = {0, 0, 0, 0};
            if (lmax < WIDE_FROM) {
                uint32_t s32[4];
                check(cudaMemcpy(s32, lo[lmax].sum, 16, cudaMemcpyDeviceToHost), "cudaMemcpy");
                for (int i = 0; i < 4; ++i) sums64[i] = s32[i];
            } else {
                check(cudaMemcpy(sums64, hi[lmax].sum, 32, cudaMemcpyDeviceToHost), "cudaMemcpy");
            }
            const uint64_t n = nx * ny * nz;
            uint32_t color = 0;
            for (int i = 0; i < 4; ++i) color |= (uint32_t)(sums64[i] / n) << (8 * i);
            const uint32_t rec[10] = {0, 0, 0, 0, 0, 0, 0, 0, color, LEAF_BIT};
            check(cudaMemcpy(d_nodes, rec, 40, cudaMemcpyHostToDevice), "cudaMemcpy");
        } else {
            // 2. top-down: root gets index 0
            const uint32_t zero = 0;
            check(cudaMemcpyAsync(root_si, &zero, 4, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync");
            for (int L = lmax; L >= 2; --L) {
                const uint32_t extent = 1u << L, depth = (uint32_t)(lmax - L);
                const uint32_t cells = (uint32_t)(dim >> L);
                const int blocks = blocks_for((uint64_t)cells * cells * cells);
                uint32_t* used = d_used + depth + 1;
                if (L < WIDE_FROM)
                    emit_level<uint32_t, uint32_t><<<blocks, 256, 0, stream>>>(lo[L], lo[L - 1], d, extent, depth, d_nodes, used);
                else if (L == WIDE_FROM)
                    emit_level<uint32_t, uint64_t><<<blocks, 256, 0, stream>>>(hi[L], lo[L - 1], d, extent, depth, d_nodes, used);
                else
                    emit_level<uint64_t, uint64_t><<<blocks, 256, 0, stream>>>(hi[L], hi[L - 1], d, extent, depth, d_nodes, used);
            }
            {
                const uint32_t cells = lo[1].cells;
                emit_voxels<<<blocks_for((uint64_t)cells * cells * cells), 256, 0, stream>>>(lo[1], d_grid, d, (uint32_t)(lmax - 1),
                                                                                         d_nodes, d_used + lmax);
            }
            check(cudaGetLastError(), "emit kernels");
        }
        // 3. ropes
        if (rope) {
            if (count == 1) {
                // single leaf: every neighbour is outside the cube -> 0 (already zero)
            } else {
                for (int L = lmax; L >= 1; --L) {
                    const uint32_t cells = (uint32_t)(dim >> L);
                    const uint32_t* si = L < WIDE_FROM ? lo[L].si : hi[L].si;
                    rope_level<<<blocks_for((uint64_t)cells * cells * cells), 256, 0, stream>>>(
                        si, cells, 1u << L, (uint32_t)(lmax - L), dim, d_nodes);
                }
                check(cudaGetLastError(), "rope kernels");
            }
        }
        std::vector<uint32_t> used(lmax + 2, 0);
        check(cudaMemcpyAsync(used.data(), d_used, (lmax + 2) * 4, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
        check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
        uint64_t depth = 0;
        for (int dpt = 0; dpt <= lmax + 1; ++dpt)
            if (used[dpt]) depth = (uint64_t)dpt;
        stats.total_nodes = count;
        stats.total_leaves = count - (count - 1) / 8; // every interior node has exactly 8 children
        stats.unique_leaves = stats.total_leaves;
        stats.depth = depth;
    } catch (...) {
        cudaFree(d_nodes);
        throw;
    }
    *d_nodes_out = d_nodes;
    *count_out = count;
    if (stats_out) *stats_out = stats;
}

} // namespace xn
