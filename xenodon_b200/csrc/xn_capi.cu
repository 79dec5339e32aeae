// xn_capi.cu -- implementation of the C ABI declared in include/xenodon_b200.h.
// One xn_ctx = one CUDA device + one stream + one offscreen RGBA8 target + one resident
// volume (grid and/or octree).  No CPU fallback: every compute entry point fails with
// XN_ERR_CUDA when the CUDA runtime cannot provide a device.
#include <cuda_runtime.h>

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "host/xn_host.hpp"
#include "xenodon_b200.h"
#include "xn_convert.h"
#include "xn_kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}

struct CudaError {
    cudaError_t e;
    const char* what;
};
#define XN_CUDA(call)                                \
    do {                                             \
        cudaError_t e__ = (call);                    \
        if (e__ != cudaSuccess) throw CudaError{e__, #call}; \
    } while (0)

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return XN_OK;
    } catch (const xn::Error& e) {
        return fail(e.status, e.what());
    } catch (const CudaError& e) {
        cudaGetLastError(); // clear sticky-less errors
        return fail(XN_ERR_CUDA, std::string(e.what) + ": " + cudaGetErrorString(e.e));
    } catch (const std::bad_alloc&) {
        return fail(XN_ERR_LIMIT, "out of host memory");
    } catch (const std::exception& e) {
        return fail(XN_ERR_INVALID, e.what());
    }
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) XN_CUDA(cudaSetDevice(dev));
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

const char* const TRAVERSAL_NAMES[5] = {"dda", "svo-naive", "esvo", "svo-df", "svo-rope"};

} // namespace

struct xn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t ev_mark[2] = {nullptr, nullptr};
    cudaEvent_t ev_gather = nullptr; // "this context's frame is rendered", waited on by xn_frame_gather
    bool timing_pending = false;
    uint64_t launches = 0;

    // pipelined frame output: two alternating targets + a copy stream
    cudaStream_t copy_stream = nullptr;
    // pipelined frame output: PIPE_DEPTH device targets in rotation, so the traversal may run up to
    // PIPE_DEPTH - 1 frames ahead of the copy-out (two balanced stages with only two buffers stall on
    // each other's jitter: 8 GPUs feeding one host at its ingest limit, profiles/README.md)
    static constexpr int PIPE_DEPTH = 3;
    uint32_t* pipe_target[PIPE_DEPTH] = {};
    uint64_t pipe_target_px = 0;
    cudaEvent_t ev_rendered[PIPE_DEPTH] = {}, ev_copied[PIPE_DEPTH] = {};
    bool copy_in_flight[PIPE_DEPTH] = {};
    bool frame_read_in_flight = false;
    int pipe_next = 0;
    double last_ms = 0;

    // volume
    uint32_t* grid = nullptr;   // x-major linear copy (may be absent while the bricked copy is resident)
    uint32_t* bricks = nullptr; // bricked copy (xn_brick.h), what the DDA reads when present
    xn::BrickLayout brick_layout{};
    cudaArray_t tex_array = nullptr; // texture residency: 3-D array + two views of it
    cudaTextureObject_t tex_unorm = 0, tex_raw = 0;
    int layout_mode = XN_GRID_LAYOUT_AUTO;
    uint64_t nx = 0, ny = 0, nz = 0;
    xn::DNode* nodes = nullptr;  // file-order 64-byte records (svo_rope on trees beyond the RNode limits)
    xn::RNode* rnodes = nullptr; // file-order 32-byte rope records (svo_rope); exactly one of the two is resident
    xn::CNode* cnodes = nullptr; // compact level-order records of the internal nodes (the other three)
    uint32_t* top_table = nullptr; // svo_naive entry table
    uint32_t top_levels = 0;
    uint64_t node_count = 0, internal_count = 0, side = 0;
    uint32_t brick_base = 0xFFFFFFFFu; // word offset of the first leaf-brick record in cnodes
    uint32_t root_meta = 0, max_depth = 0;
    bool grid_has_black_background = false;
    uint4* skip_table = nullptr; // DDA skip table of the resident grid (any layout)
    uint32_t skip_dim[3] = {0, 0, 0}, skip_shift = 0;
    double skip_uniform = 0.0; // share of the grid's bricks that hold one colour
    uint32_t* ray_pool = nullptr; // counter of the persistent ray pool (xn_kernels.cu, XN_RAY_POOL=1)

    // target
    xn_rect output{0, 0, 0, 0}, display{0, 0, 0, 0};
    bool have_target = false;
    uint32_t* own_target = nullptr;
    uint64_t own_target_px = 0;
    uint32_t* ext_target = nullptr;
    uint64_t ext_stride = 0;

    // params
    float ratio[3] = {1, 1, 1};
    uint32_t model_dim[3] = {0, 0, 0};
    float emission = 1.0f;
    bool have_params = false;

    uint32_t il_count = 1, il_index = 0;
    bool strict = false;

    std::vector<void*> ipc_opened;

    void free_texture() {
        if (tex_unorm) cudaDestroyTextureObject(tex_unorm);
        if (tex_raw) cudaDestroyTextureObject(tex_raw);
        if (tex_array) cudaFreeArray(tex_array);
        tex_unorm = tex_raw = 0;
        tex_array = nullptr;
    }
    void free_grid() {
        if (grid) cudaFree(grid);
        if (bricks) cudaFree(bricks);
        free_texture();
        if (skip_table) cudaFree(skip_table);
        skip_table = nullptr;
        grid = bricks = nullptr;
        nx = ny = nz = 0;
    }
    bool have_grid() const { return grid || bricks || tex_array; }
    bool have_svo() const { return nodes || rnodes; }
    void free_nodes() {
        if (nodes) cudaFree(nodes);
        if (rnodes) cudaFree(rnodes);
        rnodes = nullptr;
        if (cnodes) cudaFree(cnodes);
        if (top_table) cudaFree(top_table);
        nodes = nullptr;
        cnodes = nullptr;
        top_table = nullptr;
        node_count = internal_count = side = 0;
        brick_base = 0xFFFFFFFFu;
    }
};

namespace {

void check_ctx(const xn_ctx* ctx) {
    if (!ctx) throw xn::Error(XN_ERR_INVALID, "null context");
}

void fill_params(xn_ctx* ctx, int traversal, const float fwd[3], const float up[3], const float tr[3],
                 xn::FrameParams& p) {
    if (traversal < 0 || traversal > 4) throw xn::Error(XN_ERR_INVALID, "Invalid shader");
    if (!ctx->have_target) throw xn::Error(XN_ERR_INVALID, "xn_set_target has not been called");
    if (!ctx->have_params) throw xn::Error(XN_ERR_INVALID, "xn_set_params has not been called");
    if (!fwd || !up || !tr) throw xn::Error(XN_ERR_INVALID, "null camera vector");
    if (traversal == XN_DDA) {
        if (!ctx->have_grid())
            throw xn::Error(XN_ERR_INVALID, "Shader 'dda' is incompatible with model type 'svo' (requires 'tiff')");
    } else if (!ctx->have_svo()) {
        throw xn::Error(XN_ERR_INVALID, std::string("Shader '") + TRAVERSAL_NAMES[traversal] +
                                            "' is incompatible with model type 'tiff' (requires 'svo')");
    }
    std::memset(&p, 0, sizeof p);
    for (int i = 0; i < 3; ++i) {
        p.fwd[i] = fwd[i];
        p.up[i] = up[i];
        p.pos[i] = tr[i] / ctx->ratio[i]; // pre-divide, src/render/Renderer.cpp:62 (binary32)
        p.ratio[i] = ctx->ratio[i];
        p.model_dim[i] = ctx->model_dim[i];
    }
    p.out_x = ctx->output.x;
    p.out_y = ctx->output.y;
    p.out_w = ctx->output.w;
    p.out_h = ctx->output.h;
    p.disp_x = ctx->display.x;
    p.disp_y = ctx->display.y;
    p.disp_w = ctx->display.w;
    p.disp_h = ctx->display.h;
    p.emission = ctx->emission;
    if (ctx->ext_target) {
        p.target = ctx->ext_target;
        p.target_stride = ctx->ext_stride;
    } else {
        p.target = ctx->own_target;
        p.target_stride = ctx->output.w;
    }
    p.grid = ctx->grid;
    if (ctx->bricks) {
        const xn::BrickLayout& L = ctx->brick_layout;
        p.grid = ctx->bricks;
        for (int i = 0; i < 3; ++i) {
            p.bk_mask[i] = (uint32_t)L.mask[i];
            p.bk_hs[i] = L.hs[i];
        }
        p.bk_top = L.top;
        p.bk_mask_z64 = L.mask[2];
        p.bk_slots = L.total;
    }
    p.tex_unorm = ctx->tex_unorm;
    p.tex_raw = ctx->tex_raw;
    p.skip_table = ctx->skip_table;
    for (int i = 0; i < 3; ++i) p.skip_dim[i] = ctx->skip_dim[i];
    p.skip_shift = ctx->skip_shift;
    p.nx = (uint32_t)ctx->nx;
    p.ny = (uint32_t)ctx->ny;
    p.nz = (uint32_t)ctx->nz;
    p.nodes = ctx->nodes;
    p.rnodes = ctx->rnodes;
    p.cnodes = ctx->cnodes;
    {
        // XN_ESVO_BRICKS=0: the fast-mode ESVO descends into leaf bricks child by child, as the strict mode does
        // XN_ESVO_BRICKS=1: the fast-mode ESVO integrates leaf bricks in closed form (esvo_leaf_brick;
        // measured: +4 % on the dense bunny tree, -20 % on the sparse 2048^3 tree, hence an opt-in)
        const char* e = std::getenv("XN_ESVO_BRICKS");
        const bool on = e && e[0] == '1' && ctx->brick_base != 0xFFFFFFFFu;
        p.brick_base = on ? ctx->brick_base : 0xFFFFFFFFu;
    }
    p.top_table = ctx->top_table;
    p.top_levels = ctx->top_levels;
    p.root_meta = ctx->root_meta;
    p.max_depth = ctx->max_depth;
    p.il_count = ctx->il_count;
    p.il_index = ctx->il_index;
    p.skip_empty = ctx->grid_has_black_background ? 1u : 0u;
    p.pool = ctx->ray_pool;
}

// after a grid became resident: does it have a black background worth skipping (>= 25 % black)?
void classify_grid(xn_ctx* ctx) {
    const uint64_t n = ctx->nx * ctx->ny * ctx->nz;
    unsigned long long* d = nullptr;
    XN_CUDA(cudaMalloc(&d, 16));
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemsetAsync(d, 0, 16, ctx->stream);
    if (e == cudaSuccess) e = xn::launch_count_black(ctx->grid, n, n > (1ull << 24) ? 61 : 1, d, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) throw CudaError{e, "classify_grid"};
    ctx->grid_has_black_background = h[1] > 0 && h[0] * 4 >= h[1];

    // DDA skip table (uniform bricks + how far their colour extends), from the linear copy that
    // every upload path produces first.  XN_DDA_SKIP=0 leaves it out (A/B runs);
    // XN_SKIP_SHIFT = log2 of the brick edge (default 3), XN_SKIP_CAP = largest radius in bricks.
    if (ctx->skip_table) cudaFree(ctx->skip_table);
    ctx->skip_table = nullptr;
    const char* on = std::getenv("XN_DDA_SKIP");
    if (on && on[0] == '0') return;
    // brick edge: 8 voxels; 4 on the volumes far beyond L2 (a finer table finds more uniform bricks
    // between the filaments of the gas volumes: cfg3 +7 %, cfg4 +2 %; on the bunny shape it costs 3 %)
    uint32_t shift = n >= (1ull << 29) ? 2 : 3, cap = 32;
    if (const char* s = std::getenv("XN_SKIP_SHIFT")) shift = (uint32_t)std::strtoul(s, nullptr, 10);
    if (const char* s = std::getenv("XN_SKIP_CAP")) cap = (uint32_t)std::strtoul(s, nullptr, 10);
    if (ctx->nx > 0x7FFFFFu || ctx->ny > 0x7FFFFFu || ctx->nz > 0x7FFFFFu) return; // texel centres exact below 2^23
    e = xn::build_skip_table(ctx->grid, (uint32_t)ctx->nx, (uint32_t)ctx->ny, (uint32_t)ctx->nz, shift, cap,
                             &ctx->skip_table, ctx->skip_dim, &ctx->skip_uniform, ctx->stream);
    if (e != cudaSuccess) {
        // the table only saves fetches: without it the march reads every texel
        cudaGetLastError();
        ctx->skip_table = nullptr;
        return;
    }
    ctx->skip_shift = shift;
}

// ---- resident layout of the grid (xn_brick.h) ----

bool env_force_idx64() {
    const char* e = std::getenv("XN_FORCE_IDX64");
    return e && e[0] == '1';
}

// layout the bricked copy of this grid would have; false if the DDA has no cursor for it
bool plan_bricks(const xn_ctx* ctx, xn::BrickLayout& L) {
    L = xn::make_brick_layout(ctx->nx, ctx->ny, ctx->nz, env_force_idx64() ? 2 : -1);
    if (L.total <= (1ull << 32) && !env_force_idx64()) return true;
    if (L.top != 2) L = xn::make_brick_layout(ctx->nx, ctx->ny, ctx->nz, 2);
    return L.hs[2] <= 32;
}

// which residency the layout mode asks for.  AUTO: volumes far larger than L2 go to the texture
// residency, where a warp's texels share sectors whatever the ray direction and the texture unit
// does the addressing (measured: profiles/README.md); small grids stay x-major linear.
int wanted_layout(const xn_ctx* ctx, xn::BrickLayout& L) {
    const uint64_t voxels = ctx->nx * ctx->ny * ctx->nz;
    const bool tex_ok = ctx->nx <= 16384 && ctx->ny <= 16384 && ctx->nz <= 16384;
    switch (ctx->layout_mode) {
        case XN_GRID_LAYOUT_LINEAR: return XN_GRID_LAYOUT_LINEAR;
        case XN_GRID_LAYOUT_BRICKED: return plan_bricks(ctx, L) ? XN_GRID_LAYOUT_BRICKED : XN_GRID_LAYOUT_LINEAR;
        case XN_GRID_LAYOUT_TEXTURE: return tex_ok ? XN_GRID_LAYOUT_TEXTURE : XN_GRID_LAYOUT_LINEAR;
        default: break;
    }
    uint64_t min_voxels = 1ull << 29;
    if (const char* e = std::getenv("XN_TEXTURE_MIN_VOXELS")) min_voxels = std::strtoull(e, nullptr, 10);
    if (tex_ok && voxels >= min_voxels) return XN_GRID_LAYOUT_TEXTURE;
    // Smaller volumes are bound by the texture unit's rate there (one quad per clock per SM) and stay
    // linear -- unless the skip table makes most fetches unnecessary: with half of the bricks
    // uniform the texture-path march with the table wins (bunny-shape 512x361x512: 6766 vs 5595
    // Mrays/s on config 1, 3038 vs 2216 on the config 2 path)
    if (tex_ok && ctx->skip_table && ctx->skip_uniform >= 0.5 && voxels >= (1ull << 24)) return XN_GRID_LAYOUT_TEXTURE;
    return XN_GRID_LAYOUT_LINEAR;
}

cudaMemcpy3DParms array_copy_parms(xn_ctx* ctx, bool to_array) {
    cudaMemcpy3DParms cp{};
    const cudaPitchedPtr lin = make_cudaPitchedPtr(ctx->grid, ctx->nx * 4, ctx->nx, ctx->ny);
    if (to_array) {
        cp.srcPtr = lin;
        cp.dstArray = ctx->tex_array;
    } else {
        cp.srcArray = ctx->tex_array;
        cp.dstPtr = lin;
    }
    cp.extent = make_cudaExtent(ctx->nx, ctx->ny, ctx->nz);
    cp.kind = cudaMemcpyDeviceToDevice;
    return cp;
}

void ensure_linear(xn_ctx* ctx) {
    if (ctx->grid) return;
    if (!ctx->bricks && !ctx->tex_array) return;
    XN_CUDA(cudaMalloc(&ctx->grid, ctx->nx * ctx->ny * ctx->nz * 4));
    if (ctx->bricks) {
        XN_CUDA(xn::launch_unbrick_grid(ctx->bricks, ctx->grid, ctx->brick_layout, (uint32_t)ctx->nx, (uint32_t)ctx->ny,
                                        (uint32_t)ctx->nz, ctx->stream));
    } else {
        const cudaMemcpy3DParms cp = array_copy_parms(ctx, false);
        XN_CUDA(cudaMemcpy3DAsync(&cp, ctx->stream));
    }
    XN_CUDA(cudaStreamSynchronize(ctx->stream));
}

void make_texture_unguarded(xn_ctx* ctx);
void make_texture(xn_ctx* ctx) {
    try {
        make_texture_unguarded(ctx);
    } catch (...) {
        ctx->free_texture(); // never leave a half-built residency behind
        throw;
    }
}
void make_texture_unguarded(xn_ctx* ctx) {
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    XN_CUDA(cudaMalloc3DArray(&ctx->tex_array, &fmt, make_cudaExtent(ctx->nx, ctx->ny, ctx->nz)));
    const cudaMemcpy3DParms cp = array_copy_parms(ctx, true);
    XN_CUDA(cudaMemcpy3DAsync(&cp, ctx->stream));
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = ctx->tex_array;
    cudaTextureDesc td{};
    // nearest, unnormalised coordinates, transparent-black border: the reference's sampler
    // (src/render/DdaRaytraceAlgorithm.cpp:16-34)
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint;
    td.normalizedCoords = 0;
    td.readMode = cudaReadModeNormalizedFloat;
    XN_CUDA(cudaCreateTextureObject(&ctx->tex_unorm, &res, &td, nullptr));
    td.readMode = cudaReadModeElementType;
    XN_CUDA(cudaCreateTextureObject(&ctx->tex_raw, &res, &td, nullptr));
    XN_CUDA(cudaStreamSynchronize(ctx->stream));
}

// make the resident copies match the layout mode: exactly one copy stays resident
void apply_layout(xn_ctx* ctx) {
    if (!ctx->have_grid()) return;
    xn::BrickLayout L;
    const int want = wanted_layout(ctx, L);
    const int have = ctx->tex_array ? XN_GRID_LAYOUT_TEXTURE : (ctx->bricks ? XN_GRID_LAYOUT_BRICKED : XN_GRID_LAYOUT_LINEAR);
    if (want != have || (want != XN_GRID_LAYOUT_LINEAR && ctx->grid)) {
        if (want != have) {
            ensure_linear(ctx);
            if (ctx->bricks) cudaFree(ctx->bricks);
            ctx->bricks = nullptr;
            ctx->free_texture();
            if (want == XN_GRID_LAYOUT_BRICKED) {
                XN_CUDA(cudaMalloc(&ctx->bricks, L.total * 4));
                ctx->brick_layout = L;
                XN_CUDA(xn::launch_brick_grid(ctx->grid, ctx->bricks, L, (uint32_t)ctx->nx, (uint32_t)ctx->ny,
                                              (uint32_t)ctx->nz, ctx->stream));
                XN_CUDA(cudaStreamSynchronize(ctx->stream));
            } else if (want == XN_GRID_LAYOUT_TEXTURE) {
                try {
                    make_texture(ctx);
                } catch (const CudaError&) {
                    // AUTO only optimises: when the array does not fit beside the linear copy, the
                    // linear copy stays in use (an explicit request for the texture residency fails)
                    if (ctx->layout_mode != XN_GRID_LAYOUT_AUTO) throw;
                    cudaGetLastError();
                    return;
                }
            }
        }
        if (want != XN_GRID_LAYOUT_LINEAR) {
            if (ctx->grid) cudaFree(ctx->grid);
            ctx->grid = nullptr;
        }
    }
}

// L2 access-policy window over the head of the compact node array.  The records are in level
// order, so the head IS the upper levels of the tree, which every ray walks: marking them
// persisting keeps the streaming lower levels (and the frame stores) from evicting them.
// XN_L2_WINDOW_MB: size of the window (default 0 = off: measured without effect, profiles/README.md); the device's persisting-L2 limit
// is raised to match.  Trees that fit in the window entirely gain nothing (they live in L2 anyway).
void apply_l2_window(xn_ctx* ctx) {
    uint64_t window_mb = 0;
    if (const char* e = std::getenv("XN_L2_WINDOW_MB")) window_mb = std::strtoull(e, nullptr, 10);
    cudaStreamAttrValue attr{};
    const uint64_t bytes = ctx->internal_count * sizeof(xn::CNode);
    if (window_mb == 0 || !ctx->cnodes || bytes <= (window_mb << 20)) {
        attr.accessPolicyWindow.num_bytes = 0; // no window
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
        return;
    }
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    uint64_t want = window_mb << 20;
    want = std::min<uint64_t>(want, (uint64_t)std::max(max_persist, 0));
    want = std::min<uint64_t>(want, (uint64_t)std::max(max_window, 0));
    if (want == 0) return;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    attr.accessPolicyWindow.base_ptr = ctx->cnodes;
    attr.accessPolicyWindow.num_bytes = want;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

void finish_svo_upload(xn_ctx* ctx, void* d_raw, uint64_t count, uint64_t side) {
    // d_raw: count 40-byte nodes on the device; produce the 64-byte resident layout
    ctx->free_nodes();
    uint32_t* d_max = nullptr;
    XN_CUDA(cudaMalloc(&d_max, sizeof(uint32_t)));
    XN_CUDA(cudaMemsetAsync(d_max, 0, sizeof(uint32_t), ctx->stream));
    XN_CUDA(xn::launch_max_depth(d_raw, count, d_max, ctx->stream));
    uint32_t maxd = 0;
    XN_CUDA(cudaMemcpyAsync(&maxd, d_max, 4, cudaMemcpyDeviceToHost, ctx->stream));
    XN_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_max);
    if (maxd > 23) throw xn::Error(XN_ERR_LIMIT, "octree deeper than 23 levels (the traversal stack depth of the reference)");
    // svo_rope reads 32-byte records when the tree fits their fields (index 28 bits, depth 4 bits),
    // else the 64-byte ones; XN_ROPE_RECORDS=64 forces the latter (A/B runs, tests of both)
    const char* rr = std::getenv("XN_ROPE_RECORDS");
    const bool small_records = count < xn::RNODE_MAX_NODES && maxd <= xn::RNODE_MAX_DEPTH && !(rr && std::strcmp(rr, "64") == 0);
    if (small_records) {
        XN_CUDA(cudaMalloc(&ctx->rnodes, count * sizeof(xn::RNode)));
        XN_CUDA(xn::launch_relayout_rnodes(d_raw, count, ctx->rnodes, ctx->stream));
    } else {
        XN_CUDA(cudaMalloc(&ctx->nodes, count * sizeof(xn::DNode)));
        XN_CUDA(xn::launch_relayout(d_raw, count, ctx->nodes, ctx->stream));
    }
    {
        const cudaError_t e = xn::build_compact_nodes(d_raw, count, &ctx->cnodes, &ctx->internal_count, &ctx->brick_base, ctx->stream);
        if (e == cudaErrorInvalidValue) {
            cudaGetLastError();
            ctx->free_nodes();
            throw xn::Error(XN_ERR_LIMIT, "octree has more than 2^28 internal nodes");
        }
        XN_CUDA(e);
    }
    uint32_t root[2] = {0, 0};
    XN_CUDA(cudaMemcpyAsync(root, (const uint8_t*)d_raw + 32, 8, cudaMemcpyDeviceToHost, ctx->stream));
    XN_CUDA(cudaStreamSynchronize(ctx->stream));
    apply_l2_window(ctx);
    ctx->root_meta = xn::make_meta(root[0], root[1]);
    // svo_naive entry table over the first min(tree depth, TOP_LEVELS_MAX) levels
    ctx->top_levels = std::min<uint32_t>(maxd, xn::TOP_LEVELS_MAX);
    if (const char* e = std::getenv("XN_TOP_LEVELS")) // tuning knob
        ctx->top_levels = std::min<uint32_t>({(uint32_t)std::strtoul(e, nullptr, 10), maxd, 9u});
    XN_CUDA(cudaMalloc(&ctx->top_table, sizeof(uint32_t) << (3 * ctx->top_levels)));
    XN_CUDA(xn::launch_top_table(ctx->cnodes, ctx->root_meta, ctx->top_levels, ctx->top_table, ctx->stream));
    XN_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->max_depth = maxd;
    ctx->node_count = count;
    ctx->side = side;
}

} // namespace

extern "C" {

const char* xn_last_error(void) { return g_last_error.c_str(); }
const char* xn_version(void) { return "xenodon-b200 0.1 (sm_100a)"; }

int xn_traversal_from_name(const char* name) {
    if (!name) return XN_ERR_INVALID;
    for (int i = 0; i < 5; ++i)
        if (std::strcmp(name, TRAVERSAL_NAMES[i]) == 0) return i;
    return fail(XN_ERR_INVALID, std::string("Invalid shader '") + name + "'");
}
const char* xn_traversal_name(int t) { return t >= 0 && t < 5 ? TRAVERSAL_NAMES[t] : "?"; }

int xn_device_count(int* count) {
    return guarded([&] {
        if (!count) throw xn::Error(XN_ERR_INVALID, "null argument");
        *count = 0;
        XN_CUDA(cudaGetDeviceCount(count));
    });
}

int xn_device_name(int device, char* buf, size_t cap) {
    return guarded([&] {
        if (!buf || cap == 0) throw xn::Error(XN_ERR_INVALID, "null argument");
        cudaDeviceProp prop;
        XN_CUDA(cudaGetDeviceProperties(&prop, device));
        std::snprintf(buf, cap, "%s", prop.name);
    });
}

int xn_ctx_create(int cuda_device, xn_ctx** out) {
    return guarded([&] {
        if (!out) throw xn::Error(XN_ERR_INVALID, "null argument");
        *out = nullptr;
        int n = 0;
        XN_CUDA(cudaGetDeviceCount(&n));
        if (cuda_device < 0 || cuda_device >= n)
            throw xn::Error(XN_ERR_INVALID, "device index " + std::to_string(cuda_device) + " out of range (" +
                                                std::to_string(n) + " CUDA devices)");
        DeviceGuard g(cuda_device);
        auto ctx = std::make_unique<xn_ctx>();
        ctx->device = cuda_device;
        if (const char* e = std::getenv("XN_GRID_LAYOUT")) { // default mode of new contexts (A/B runs)
            if (std::strcmp(e, "linear") == 0) ctx->layout_mode = XN_GRID_LAYOUT_LINEAR;
            else if (std::strcmp(e, "bricked") == 0) ctx->layout_mode = XN_GRID_LAYOUT_BRICKED;
            else if (std::strcmp(e, "texture") == 0) ctx->layout_mode = XN_GRID_LAYOUT_TEXTURE;
        }
        XN_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        XN_CUDA(cudaEventCreate(&ctx->ev_start));
        XN_CUDA(cudaEventCreate(&ctx->ev_stop));
        XN_CUDA(cudaEventCreate(&ctx->ev_mark[0]));
        XN_CUDA(cudaEventCreate(&ctx->ev_mark[1]));
        XN_CUDA(cudaEventCreateWithFlags(&ctx->ev_gather, cudaEventDisableTiming));
        XN_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        XN_CUDA(cudaMalloc(&ctx->ray_pool, 256));
        for (int i = 0; i < xn_ctx::PIPE_DEPTH; ++i) {
            XN_CUDA(cudaEventCreateWithFlags(&ctx->ev_rendered[i], cudaEventDisableTiming));
            XN_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
        }
        *out = ctx.release();
    });
}

int xn_ctx_destroy(xn_ctx* ctx) {
    return guarded([&] {
        if (!ctx) return;
        DeviceGuard g(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 2; ++i)
            if (ctx->ev_mark[i]) cudaEventDestroy(ctx->ev_mark[i]);
        for (int i = 0; i < xn_ctx::PIPE_DEPTH; ++i) {
            if (ctx->pipe_target[i]) cudaFree(ctx->pipe_target[i]);
            if (ctx->ev_rendered[i]) cudaEventDestroy(ctx->ev_rendered[i]);
            if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        }
        if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
        for (void* p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);
        ctx->free_grid();
        ctx->free_nodes();
        if (ctx->own_target) cudaFree(ctx->own_target);
        if (ctx->ray_pool) cudaFree(ctx->ray_pool);
        cudaEventDestroy(ctx->ev_start);
        cudaEventDestroy(ctx->ev_stop);
        if (ctx->ev_gather) cudaEventDestroy(ctx->ev_gather);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
    });
}

int xn_ctx_device(const xn_ctx* ctx) { return ctx ? ctx->device : XN_ERR_INVALID; }

// ---- volumes ----

static void check_grid_dims(uint64_t nx, uint64_t ny, uint64_t nz) {
    if (nx == 0 || ny == 0 || nz == 0) throw xn::Error(XN_ERR_INVALID, "empty grid");
    if (nx > 0xFFFFFFFFull || ny > 0xFFFFFFFFull || nz > 0xFFFFFFFFull)
        throw xn::Error(XN_ERR_LIMIT, "grid dimension exceeds 2^32 - 1");
}

int xn_upload_grid(xn_ctx* ctx, const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz) {
    return guarded([&] {
        check_ctx(ctx);
        if (!rgba) throw xn::Error(XN_ERR_INVALID, "null grid");
        check_grid_dims(nx, ny, nz);
        DeviceGuard g(ctx->device);
        ctx->free_grid();
        const uint64_t bytes = nx * ny * nz * 4;
        XN_CUDA(cudaMalloc(&ctx->grid, bytes));
        // bulk copy in 256 MiB pieces (pageable source; replaces the reference's scalar
        // element-wise staging loop, src/render/DdaRaytraceAlgorithm.cpp:61-63)
        const uint64_t piece = 256ull << 20;
        for (uint64_t off = 0; off < bytes; off += piece)
            XN_CUDA(cudaMemcpyAsync((uint8_t*)ctx->grid + off, rgba + off, std::min(piece, bytes - off),
                                    cudaMemcpyHostToDevice, ctx->stream));
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->nx = nx;
        ctx->ny = ny;
        ctx->nz = nz;
        classify_grid(ctx);
        apply_layout(ctx);
    });
}

// Volume ingest pipeline: replaces TIFFReadRGBAImage per directory + the scalar
// staging loop + the blocking upload of the reference (src/model/Grid.cpp:60-75,
// src/render/DdaRaytraceAlgorithm.cpp:49-96).  Worker threads pread the raw sample bytes of whole
// z slices into page-locked buffers and hand them to the device on their own streams, where a
// kernel does the decoding (sample expansion, alpha pre-multiplication, row flip) straight into the
// resident grid: no host-side per-voxel work, no host copy of the volume, disk reads of one slice
// overlap the transfer and decode of others.
int xn_upload_grid_tiff(xn_ctx* ctx, const char* path, uint64_t dims_out[3], double* seconds_out) {
    return guarded([&] {
        check_ctx(ctx);
        if (!path) throw xn::Error(XN_ERR_INVALID, "null path");
        const auto t_begin = std::chrono::steady_clock::now();
        const xn::TiffPlan plan = xn::tiff_plan(path);
        const uint64_t nx = plan.info.nx, ny = plan.info.ny, nz = plan.info.nz;
        check_grid_dims(nx, ny, nz);
        if (dims_out) dims_out[0] = nx, dims_out[1] = ny, dims_out[2] = nz;
        if (!plan.streamable) {
            // tiled / mixed-format files: decode on the host as before, then one bulk upload
            std::vector<uint8_t> host(nx * ny * nz * 4);
            xn::tiff_read(path, host.data(), host.size());
            const int rc = xn_upload_grid(ctx, host.data(), nx, ny, nz);
            if (rc != XN_OK) throw xn::Error(rc, g_last_error);
            if (seconds_out)
                *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
            return;
        }
        DeviceGuard g(ctx->device);
        ctx->free_grid();
        const uint64_t layer_px = nx * ny, raw_bytes = layer_px * plan.format.samples;
        XN_CUDA(cudaMalloc(&ctx->grid, layer_px * nz * 4));
        const xn::TiffDecode fmt{plan.format.samples, plan.format.photometric, plan.format.has_alpha,
                                 plan.format.unassociated, plan.format.flip};
        unsigned want = std::thread::hardware_concurrency();
        if (const char* e = std::getenv("XN_INGEST_THREADS")) want = (unsigned)std::strtoul(e, nullptr, 10);
        const unsigned n_workers = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)want, 8ull, nz}));
        std::atomic<uint64_t> next{0};
        std::mutex err_mutex;
        std::string err_msg;
        int err_status = XN_OK;
        auto worker = [&](unsigned) {
            uint8_t* pinned = nullptr;
            uint8_t* d_raw = nullptr;
            cudaStream_t stream = nullptr;
            int fd = -1;
            try {
                XN_CUDA(cudaSetDevice(ctx->device));
                XN_CUDA(cudaHostAlloc((void**)&pinned, raw_bytes, cudaHostAllocDefault));
                XN_CUDA(cudaMalloc((void**)&d_raw, raw_bytes));
                XN_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
                fd = ::open(path, O_RDONLY);
                if (fd < 0) throw xn::Error(XN_ERR_IO, "Failed to open");
                for (;;) {
                    const uint64_t z = next.fetch_add(1);
                    if (z >= nz) break;
                    {
                        std::lock_guard<std::mutex> lock(err_mutex);
                        if (err_status != XN_OK) break;
                    }
                    uint64_t at = 0;
                    for (const xn::TiffRun& run : plan.slices[z]) {
                        uint64_t done = 0;
                        while (done < run.bytes) {
                            const ssize_t got = ::pread(fd, pinned + at + done, run.bytes - done, (off_t)(run.offset + done));
                            if (got <= 0) throw xn::Error(XN_ERR_FORMAT, "TIFF: unexpected end of file");
                            done += (uint64_t)got;
                        }
                        at += run.bytes;
                    }
                    if (at != raw_bytes) throw xn::Error(XN_ERR_FORMAT, "TIFF: strip is shorter than its rows");
                    XN_CUDA(cudaMemcpyAsync(d_raw, pinned, raw_bytes, cudaMemcpyHostToDevice, stream));
                    XN_CUDA(xn::launch_tiff_decode(d_raw, ctx->grid + z * layer_px, (uint32_t)nx, (uint32_t)ny, fmt, stream));
                    XN_CUDA(cudaStreamSynchronize(stream)); // the page-locked buffer is reused for the next slice
                }
            } catch (const xn::Error& e) {
                std::lock_guard<std::mutex> lock(err_mutex);
                if (err_status == XN_OK) err_status = e.status, err_msg = e.what();
            } catch (const CudaError& e) {
                std::lock_guard<std::mutex> lock(err_mutex);
                if (err_status == XN_OK)
                    err_status = XN_ERR_CUDA, err_msg = std::string(e.what) + ": " + cudaGetErrorString(e.e);
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lock(err_mutex);
                if (err_status == XN_OK) err_status = XN_ERR_INVALID, err_msg = e.what();
            }
            if (fd >= 0) ::close(fd);
            if (stream) cudaStreamSynchronize(stream), cudaStreamDestroy(stream);
            if (d_raw) cudaFree(d_raw);
            if (pinned) cudaFreeHost(pinned);
        };
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < n_workers; ++i) pool.emplace_back(worker, i);
        worker(0);
        for (auto& t : pool) t.join();
        if (err_status != XN_OK) {
            ctx->free_grid();
            throw xn::Error(err_status, err_msg);
        }
        ctx->nx = nx;
        ctx->ny = ny;
        ctx->nz = nz;
        classify_grid(ctx);
        apply_layout(ctx);
        if (seconds_out)
            *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    });
}

int xn_upload_grid_device(xn_ctx* ctx, const void* d_rgba, uint64_t nx, uint64_t ny, uint64_t nz) {
    return guarded([&] {
        check_ctx(ctx);
        if (!d_rgba) throw xn::Error(XN_ERR_INVALID, "null grid");
        check_grid_dims(nx, ny, nz);
        DeviceGuard g(ctx->device);
        ctx->free_grid();
        const uint64_t bytes = nx * ny * nz * 4;
        XN_CUDA(cudaMalloc(&ctx->grid, bytes));
        XN_CUDA(cudaMemcpyAsync(ctx->grid, d_rgba, bytes, cudaMemcpyDefault, ctx->stream));
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->nx = nx;
        ctx->ny = ny;
        ctx->nz = nz;
        classify_grid(ctx);
        apply_layout(ctx);
    });
}

int xn_upload_svo(xn_ctx* ctx, const xn_node* nodes, uint64_t count, uint64_t side) {
    return guarded([&] {
        check_ctx(ctx);
        if (!nodes || count == 0) throw xn::Error(XN_ERR_INVALID, "empty octree");
        if (count > 0xFFFFFFFFull) throw xn::Error(XN_ERR_LIMIT, "octree exceeds 2^32 - 1 nodes");
        DeviceGuard g(ctx->device);
        void* d_raw = nullptr;
        XN_CUDA(cudaMalloc(&d_raw, count * sizeof(xn_node)));
        try {
            const uint64_t bytes = count * sizeof(xn_node), piece = 256ull << 20;
            for (uint64_t off = 0; off < bytes; off += piece)
                XN_CUDA(cudaMemcpyAsync((uint8_t*)d_raw + off, (const uint8_t*)nodes + off,
                                        std::min(piece, bytes - off), cudaMemcpyHostToDevice, ctx->stream));
            finish_svo_upload(ctx, d_raw, count, side);
        } catch (...) {
            cudaFree(d_raw);
            throw;
        }
        cudaFree(d_raw);
    });
}

int xn_upload_svo_device(xn_ctx* ctx, const void* d_nodes40, uint64_t count, uint64_t side) {
    return guarded([&] {
        check_ctx(ctx);
        if (!d_nodes40 || count == 0) throw xn::Error(XN_ERR_INVALID, "empty octree");
        if (count > 0xFFFFFFFFull) throw xn::Error(XN_ERR_LIMIT, "octree exceeds 2^32 - 1 nodes");
        DeviceGuard g(ctx->device);
        finish_svo_upload(ctx, const_cast<void*>(d_nodes40), count, side);
    });
}

int xn_convert_resident_grid(xn_ctx* ctx, int chan_diff, int type, int bind, xn_node** nodes_out, uint64_t* count_out,
                             uint64_t* side_out, xn_build_stats* stats_out) {
    if (chan_diff < 0 || chan_diff > 255) return fail(XN_ERR_INVALID, "channel difference must be 0..255");
    return xn_convert_resident_grid_ex(ctx, 0, (double)chan_diff, type, bind, nodes_out, count_out, side_out, stats_out);
}

int xn_convert_resident_grid_ex(xn_ctx* ctx, int heuristic, double param, int type, int bind, xn_node** nodes_out,
                                uint64_t* count_out, uint64_t* side_out, xn_build_stats* stats_out) {
    return guarded([&] {
        check_ctx(ctx);
        if (!ctx->have_grid()) throw xn::Error(XN_ERR_INVALID, "no grid is resident");
        if (heuristic != 0 && heuristic != 1) throw xn::Error(XN_ERR_INVALID, "unknown split heuristic");
        if (heuristic == 0 && !(param >= 0.0 && param <= 255.0)) throw xn::Error(XN_ERR_INVALID, "channel difference must be 0..255");
        if (heuristic == 1 && !(param >= 0.0)) throw xn::Error(XN_ERR_INVALID, "standard deviation must be >= 0");
        if (type < 0 || type > 2) throw xn::Error(XN_ERR_INVALID, "unknown octree type");
        if (nodes_out) *nodes_out = nullptr;
        DeviceGuard g(ctx->device);
        ensure_linear(ctx); // the builder reads the x-major copy
        void* d_nodes = nullptr;
        uint64_t count = 0, side = 0;
        xn_build_stats stats{};
        xn::gpu_build_octree(ctx->grid, ctx->nx, ctx->ny, ctx->nz, heuristic, param, type == 2, ctx->stream, &d_nodes,
                             &count, &side, &stats);
        if (type == 1) { // --dag: one representative per class of identical subtrees
            void* d_dag = nullptr;
            uint64_t dag_count = 0, unique_leaves = 0;
            try {
                xn::gpu_dag_from_sparse(d_nodes, count, ctx->stream, &d_dag, &dag_count, &unique_leaves);
            } catch (...) {
                cudaFree(d_nodes);
                throw;
            }
            cudaFree(d_nodes);
            d_nodes = d_dag;
            count = dag_count;
            stats.unique_leaves = unique_leaves; // total_nodes / total_leaves keep counting every visit, as the reference does
        }
        if (stats_out) *stats_out = stats;
        try {
            if (nodes_out) {
                xn_node* host = (xn_node*)std::malloc(std::max<uint64_t>(count, 1) * sizeof(xn_node));
                if (!host) throw std::bad_alloc();
                cudaError_t e = cudaMemcpyAsync(host, d_nodes, count * sizeof(xn_node), cudaMemcpyDeviceToHost, ctx->stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
                if (e != cudaSuccess) {
                    std::free(host);
                    throw CudaError{e, "cudaMemcpyAsync (octree nodes)"};
                }
                *nodes_out = host;
            }
            if (bind) finish_svo_upload(ctx, d_nodes, count, side);
        } catch (...) {
            cudaFree(d_nodes);
            throw;
        }
        cudaFree(d_nodes);
        apply_layout(ctx);
        if (count_out) *count_out = count;
        if (side_out) *side_out = side;
    });
}

int xn_synth_grid_device(xn_ctx* ctx, int kind, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t seed) {
    return guarded([&] {
        check_ctx(ctx);
        if (kind < 0 || kind > 1) throw xn::Error(XN_ERR_INVALID, "unknown synthetic volume kind");
        check_grid_dims(nx, ny, nz);
        if (nx > 0xFFFFu || ny > 0xFFFFu || nz > 0xFFFFu) throw xn::Error(XN_ERR_LIMIT, "synthetic grid too large");
        DeviceGuard g(ctx->device);
        ctx->free_grid();
        XN_CUDA(cudaMalloc(&ctx->grid, nx * ny * nz * 4));
        const xn::SynthSpec spec{(uint32_t)kind, (uint32_t)nx, (uint32_t)ny, (uint32_t)nz, seed};
        XN_CUDA(xn::launch_synth(ctx->grid, spec, ctx->stream));
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->nx = nx;
        ctx->ny = ny;
        ctx->nz = nz;
        classify_grid(ctx);
        apply_layout(ctx);
    });
}

int xn_synth_grid_host(int kind, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t seed, uint8_t* rgba_out) {
    return guarded([&] {
        if (!rgba_out) throw xn::Error(XN_ERR_INVALID, "null output");
        xn::synth_grid_host(kind, nx, ny, nz, seed, rgba_out);
    });
}

int xn_download_grid(xn_ctx* ctx, uint8_t* rgba_out, uint64_t cap_bytes) {
    return guarded([&] {
        check_ctx(ctx);
        if (!ctx->have_grid()) throw xn::Error(XN_ERR_INVALID, "no grid is resident");
        const uint64_t bytes = ctx->nx * ctx->ny * ctx->nz * 4;
        if (!rgba_out || cap_bytes < bytes) throw xn::Error(XN_ERR_INVALID, "output buffer too small");
        DeviceGuard g(ctx->device);
        ensure_linear(ctx);
        XN_CUDA(cudaMemcpyAsync(rgba_out, ctx->grid, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
        apply_layout(ctx);
    });
}

int xn_set_grid_layout(xn_ctx* ctx, int mode) {
    return guarded([&] {
        check_ctx(ctx);
        if (mode < XN_GRID_LAYOUT_AUTO || mode > XN_GRID_LAYOUT_TEXTURE) throw xn::Error(XN_ERR_INVALID, "unknown grid layout");
        DeviceGuard g(ctx->device);
        ctx->layout_mode = mode;
        apply_layout(ctx);
        if (mode != XN_GRID_LAYOUT_AUTO && mode != XN_GRID_LAYOUT_LINEAR && ctx->grid)
            throw xn::Error(XN_ERR_LIMIT, "grid shape is outside the limits of the requested layout");
    });
}

int xn_grid_layout(const xn_ctx* ctx, int* layout_out, uint64_t* resident_bytes_out) {
    if (!ctx || !layout_out) return fail(XN_ERR_INVALID, "null argument");
    *layout_out = ctx->tex_array ? XN_GRID_LAYOUT_TEXTURE
                                 : (ctx->bricks ? XN_GRID_LAYOUT_BRICKED : (ctx->grid ? XN_GRID_LAYOUT_LINEAR : XN_GRID_LAYOUT_AUTO));
    if (resident_bytes_out)
        *resident_bytes_out = (ctx->bricks ? ctx->brick_layout.total * 4 : 0) +
                              ((ctx->grid ? 1 : 0) + (ctx->tex_array ? 1 : 0)) * ctx->nx * ctx->ny * ctx->nz * 4;
    return XN_OK;
}

int xn_brick_layout(uint64_t nx, uint64_t ny, uint64_t nz, int top, uint64_t desc_out[8]) {
    return guarded([&] {
        if (!desc_out || nx == 0 || ny == 0 || nz == 0 || top > 2) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        const xn::BrickLayout L = xn::make_brick_layout(nx, ny, nz, top);
        for (int a = 0; a < 3; ++a) {
            desc_out[a] = L.mask[a];
            desc_out[3 + a] = L.hs[a];
        }
        desc_out[6] = L.top;
        desc_out[7] = L.total;
    });
}

int xn_brick_indices(uint64_t nx, uint64_t ny, uint64_t nz, int top, const int32_t* xyz, uint64_t n, uint64_t* index_out) {
    return guarded([&] {
        if (!xyz || !index_out || nx == 0 || ny == 0 || nz == 0 || top > 2) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        const xn::BrickLayout L = xn::make_brick_layout(nx, ny, nz, top);
        for (uint64_t i = 0; i < n; ++i)
            index_out[i] = xn::brick_axis(L, 0, xyz[3 * i]) | xn::brick_axis(L, 1, xyz[3 * i + 1]) |
                           xn::brick_axis(L, 2, xyz[3 * i + 2]);
    });
}

// ---- target / params ----

int xn_set_target(xn_ctx* ctx, const xn_rect* output, const xn_rect* display) {
    return guarded([&] {
        check_ctx(ctx);
        if (!output || !display) throw xn::Error(XN_ERR_INVALID, "null rectangle");
        if (display->w == 0 || display->h == 0) throw xn::Error(XN_ERR_INVALID, "empty display region");
        DeviceGuard g(ctx->device);
        const uint64_t px = (uint64_t)output->w * output->h;
        if (px > ctx->own_target_px) {
            if (ctx->own_target) cudaFree(ctx->own_target);
            ctx->own_target = nullptr;
            ctx->own_target_px = 0;
            XN_CUDA(cudaMalloc(&ctx->own_target, std::max<uint64_t>(px, 1) * 4));
            ctx->own_target_px = px;
        }
        ctx->output = *output;
        ctx->display = *display;
        ctx->have_target = true;
    });
}

int xn_set_params(xn_ctx* ctx, const float voxel_ratio[3], const uint32_t model_dim[3], float emission_coeff) {
    return guarded([&] {
        check_ctx(ctx);
        if (!voxel_ratio || !model_dim) throw xn::Error(XN_ERR_INVALID, "null argument");
        for (int i = 0; i < 3; ++i) {
            if (!(voxel_ratio[i] > 0.0f)) throw xn::Error(XN_ERR_INVALID, "voxel ratio must be positive");
            ctx->ratio[i] = voxel_ratio[i];
            ctx->model_dim[i] = model_dim[i];
        }
        if (!(emission_coeff >= 0.0f)) throw xn::Error(XN_ERR_INVALID, "emission coefficient must be >= 0");
        ctx->emission = emission_coeff;
        ctx->have_params = true;
    });
}

int xn_set_precision(xn_ctx* ctx, int mode) {
    return guarded([&] {
        check_ctx(ctx);
        if (mode != XN_PRECISION_FAST && mode != XN_PRECISION_STRICT)
            throw xn::Error(XN_ERR_INVALID, "unknown precision mode");
        ctx->strict = mode == XN_PRECISION_STRICT;
    });
}

int xn_set_interleave(xn_ctx* ctx, uint32_t count, uint32_t index) {
    return guarded([&] {
        check_ctx(ctx);
        if (count == 0 || index >= count) throw xn::Error(XN_ERR_INVALID, "interleave index must be < count");
        ctx->il_count = count;
        ctx->il_index = index;
    });
}

int xn_owned_rays(const xn_ctx* ctx, uint64_t* rays) {
    if (!ctx || !rays) return fail(XN_ERR_INVALID, "null argument");
    uint64_t rows = 0;
    const uint32_t bh = xn::BLOCK_H;
    for (uint32_t s = ctx->il_index, y = s * bh; y < ctx->output.h; s += ctx->il_count, y = s * bh)
        rows += std::min<uint64_t>(bh, ctx->output.h - y);
    *rays = rows * ctx->output.w;
    return XN_OK;
}

int xn_set_target_buffer(xn_ctx* ctx, void* device_ptr, size_t stride_px) {
    return guarded([&] {
        check_ctx(ctx);
        ctx->ext_target = (uint32_t*)device_ptr;
        ctx->ext_stride = device_ptr ? stride_px : 0;
        if (device_ptr && stride_px < ctx->output.w && ctx->have_target)
            throw xn::Error(XN_ERR_INVALID, "target stride smaller than the region width");
    });
}

// ---- frames ----

int xn_render(xn_ctx* ctx, int traversal, const float forward[3], const float up[3], const float translation[3]) {
    return guarded([&] {
        check_ctx(ctx);
        xn::FrameParams p;
        fill_params(ctx, traversal, forward, up, translation, p);
        DeviceGuard g(ctx->device);
        XN_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
        XN_CUDA(xn::launch_traversal(traversal, p, false, ctx->strict, ctx->stream));
        XN_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
        ctx->timing_pending = true;
        ++ctx->launches;
    });
}

namespace {
// Pipelined frame: traversal into one of two alternating device targets, then the region's OWNED
// rows (all of them, or this context's 16-row stripes under xn_set_interleave) copied to
// host_dst[y * stride_px + x] on the copy stream.
void render_download(xn_ctx* ctx, int traversal, const float forward[3], const float up[3], const float translation[3],
                     uint32_t* host_dst, size_t stride_px) {
    check_ctx(ctx);
    if (!host_dst) throw xn::Error(XN_ERR_INVALID, "null destination");
    if (ctx->ext_target) throw xn::Error(XN_ERR_INVALID, "pipelined output cannot target an external buffer");
    xn::FrameParams p;
    fill_params(ctx, traversal, forward, up, translation, p);
    if (stride_px == 0) stride_px = p.out_w;
    if (stride_px < p.out_w) throw xn::Error(XN_ERR_INVALID, "stride smaller than the region width");
    DeviceGuard g(ctx->device);
    const uint64_t px = (uint64_t)p.out_w * p.out_h;
    if (px == 0) return;
    if (px > ctx->pipe_target_px) {
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
        XN_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        for (int i = 0; i < xn_ctx::PIPE_DEPTH; ++i) {
            if (ctx->pipe_target[i]) cudaFree(ctx->pipe_target[i]);
            ctx->pipe_target[i] = nullptr;
            ctx->copy_in_flight[i] = false;
        }
        ctx->pipe_target_px = 0;
        for (int i = 0; i < xn_ctx::PIPE_DEPTH; ++i) XN_CUDA(cudaMalloc(&ctx->pipe_target[i], px * 4));
        ctx->pipe_target_px = px;
    }
    const int b = ctx->pipe_next;
    ctx->pipe_next = (b + 1) % xn_ctx::PIPE_DEPTH;
    // the traversal may only overwrite target b once its previous copy-out has finished
    if (ctx->copy_in_flight[b]) XN_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
    p.target = ctx->pipe_target[b];
    p.target_stride = p.out_w;
    XN_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
    XN_CUDA(xn::launch_traversal(traversal, p, false, ctx->strict, ctx->stream));
    XN_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
    XN_CUDA(cudaEventRecord(ctx->ev_rendered[b], ctx->stream));
    XN_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[b], 0));
    const uint32_t* src = ctx->pipe_target[b];
    if (ctx->il_count <= 1 && stride_px == p.out_w) {
        XN_CUDA(cudaMemcpyAsync(host_dst, src, px * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    } else if (ctx->il_count <= 1) {
        XN_CUDA(cudaMemcpy2DAsync(host_dst, stride_px * 4, src, (size_t)p.out_w * 4, (size_t)p.out_w * 4, p.out_h,
                                  cudaMemcpyDeviceToHost, ctx->copy_stream));
    } else {
        // owned stripes: BLOCK_H rows each, il_count stripes apart.  With tight rows a stripe is one
        // contiguous run, so all full stripes go in ONE 2-D copy (row = a stripe); a ragged last
        // stripe (frame height not a multiple of 16) follows on its own.
        const uint32_t bh = xn::BLOCK_H;
        const uint32_t stripes = (p.out_h + bh - 1) / bh;
        const uint32_t first = ctx->il_index;
        uint32_t full = 0, last_rows = 0, last_stripe = 0;
        for (uint32_t s = first; s < stripes; s += ctx->il_count) {
            const uint32_t rows = std::min<uint32_t>(bh, p.out_h - s * bh);
            if (rows == bh) ++full;
            else last_rows = rows, last_stripe = s;
        }
        if (stride_px == p.out_w) {
            const size_t run = (size_t)bh * p.out_w * 4, pitch = run * ctx->il_count;
            const size_t off = (size_t)first * bh * p.out_w;
            if (full)
                XN_CUDA(cudaMemcpy2DAsync(host_dst + off, pitch, src + off, pitch, run, full, cudaMemcpyDeviceToHost,
                                          ctx->copy_stream));
        } else {
            for (uint32_t s = first, k = 0; k < full; s += ctx->il_count, ++k)
                XN_CUDA(cudaMemcpy2DAsync(host_dst + (size_t)s * bh * stride_px, stride_px * 4,
                                          src + (size_t)s * bh * p.out_w, (size_t)p.out_w * 4, (size_t)p.out_w * 4, bh,
                                          cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        if (last_rows)
            XN_CUDA(cudaMemcpy2DAsync(host_dst + (size_t)last_stripe * bh * stride_px, stride_px * 4,
                                      src + (size_t)last_stripe * bh * p.out_w, (size_t)p.out_w * 4, (size_t)p.out_w * 4,
                                      last_rows, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    XN_CUDA(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
    ctx->copy_in_flight[b] = true;
    ctx->timing_pending = true;
    ++ctx->launches;
}

void CUDART_CB write_flag_cb(void* arg) {
    // arg packs nothing: the pair lives in a heap cell so that the callback can free it
    auto* cell = static_cast<std::pair<volatile uint32_t*, uint32_t>*>(arg);
    __atomic_store_n(const_cast<uint32_t*>(cell->first), cell->second, __ATOMIC_RELEASE);
    delete cell;
}
} // namespace

int xn_render_download_async(xn_ctx* ctx, int traversal, const float forward[3], const float up[3],
                             const float translation[3], uint32_t* host_dst) {
    return guarded([&] { render_download(ctx, traversal, forward, up, translation, host_dst, 0); });
}

int xn_render_download_to(xn_ctx* ctx, int traversal, const float forward[3], const float up[3],
                          const float translation[3], uint32_t* host_frame, size_t stride_px) {
    return guarded([&] { render_download(ctx, traversal, forward, up, translation, host_frame, stride_px); });
}

int xn_signal_after_copy(xn_ctx* ctx, volatile uint32_t* host_flag, uint32_t value) {
    return guarded([&] {
        check_ctx(ctx);
        if (!host_flag) throw xn::Error(XN_ERR_INVALID, "null flag");
        DeviceGuard g(ctx->device);
        auto* cell = new std::pair<volatile uint32_t*, uint32_t>(host_flag, value);
        const cudaError_t e = cudaLaunchHostFunc(ctx->copy_stream, write_flag_cb, cell);
        if (e != cudaSuccess) {
            delete cell;
            throw CudaError{e, "cudaLaunchHostFunc"};
        }
    });
}

int xn_host_register(void* p, size_t bytes) {
    return guarded([&] {
        if (!p || bytes == 0) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        XN_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    });
}

int xn_host_unregister(void* p) {
    return guarded([&] {
        if (p) XN_CUDA(cudaHostUnregister(p));
    });
}

int xn_host_alloc(size_t bytes, void** out) {
    return guarded([&] {
        if (!out) throw xn::Error(XN_ERR_INVALID, "null argument");
        *out = nullptr;
        XN_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    });
}

int xn_host_free(void* p) {
    return guarded([&] {
        if (p) XN_CUDA(cudaFreeHost(p));
    });
}

int xn_mark(xn_ctx* ctx, int which) {
    return guarded([&] {
        check_ctx(ctx);
        if (which < 0 || which > 1) throw xn::Error(XN_ERR_INVALID, "mark index must be 0 or 1");
        DeviceGuard g(ctx->device);
        XN_CUDA(cudaEventRecord(ctx->ev_mark[which], ctx->stream));
    });
}

int xn_mark_elapsed(xn_ctx* ctx, double* ms) {
    return guarded([&] {
        check_ctx(ctx);
        if (!ms) throw xn::Error(XN_ERR_INVALID, "null argument");
        DeviceGuard g(ctx->device);
        XN_CUDA(cudaEventSynchronize(ctx->ev_mark[1]));
        float f = 0;
        XN_CUDA(cudaEventElapsedTime(&f, ctx->ev_mark[0], ctx->ev_mark[1]));
        *ms = f;
    });
}

int xn_launch_count(const xn_ctx* ctx, uint64_t* count) {
    if (!ctx || !count) return fail(XN_ERR_INVALID, "null argument");
    *count = ctx->launches;
    return XN_OK;
}

int xn_sync(xn_ctx* ctx, double* kernel_ms) {
    return guarded([&] {
        check_ctx(ctx);
        DeviceGuard g(ctx->device);
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
        bool copying = ctx->frame_read_in_flight;
        for (bool f : ctx->copy_in_flight) copying = copying || f;
        if (copying) {
            XN_CUDA(cudaStreamSynchronize(ctx->copy_stream));
            for (bool& f : ctx->copy_in_flight) f = false;
            ctx->frame_read_in_flight = false;
        }
        if (ctx->timing_pending) {
            float ms = 0;
            XN_CUDA(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
            ctx->last_ms = ms;
            ctx->timing_pending = false;
        }
        if (kernel_ms) *kernel_ms = ctx->last_ms;
    });
}

int xn_download(xn_ctx* ctx, uint32_t* dst, size_t stride_px) {
    return guarded([&] {
        check_ctx(ctx);
        if (!dst) throw xn::Error(XN_ERR_INVALID, "null destination");
        if (!ctx->have_target) throw xn::Error(XN_ERR_INVALID, "xn_set_target has not been called");
        const uint32_t w = ctx->output.w, h = ctx->output.h;
        if (stride_px == 0) stride_px = w;
        if (stride_px < w) throw xn::Error(XN_ERR_INVALID, "stride smaller than the region width");
        if (w == 0 || h == 0) return;
        DeviceGuard g(ctx->device);
        const uint32_t* src = ctx->ext_target ? ctx->ext_target : ctx->own_target;
        const size_t src_stride = ctx->ext_target ? ctx->ext_stride : w;
        XN_CUDA(cudaMemcpy2DAsync(dst, stride_px * 4, src, src_stride * 4, (size_t)w * 4, h, cudaMemcpyDeviceToHost,
                                  ctx->stream));
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int xn_render_stats_pass(xn_ctx* ctx, int traversal, const float forward[3], const float up[3],
                         const float translation[3], uint32_t* steps_out, uint64_t* bytes_out,
                         uint64_t totals_out[2]) {
    return guarded([&] {
        check_ctx(ctx);
        xn::FrameParams p;
        fill_params(ctx, traversal, forward, up, translation, p);
        DeviceGuard g(ctx->device);
        const uint64_t n = (uint64_t)p.out_w * p.out_h;
        if (totals_out) totals_out[0] = totals_out[1] = 0;
        if (n == 0) return;
        uint32_t* d_steps = nullptr;
        unsigned long long *d_bytes = nullptr, *d_tot = nullptr;
        uint32_t* d_scratch_target = nullptr;
        try {
            XN_CUDA(cudaMalloc(&d_steps, n * 4));
            XN_CUDA(cudaMalloc(&d_bytes, n * 8));
            XN_CUDA(cudaMalloc(&d_tot, 16));
            // the stats pass must not disturb the image of a previous xn_render
            XN_CUDA(cudaMalloc(&d_scratch_target, n * 4));
            XN_CUDA(cudaMemsetAsync(d_tot, 0, 16, ctx->stream));
            // rows of stripes this context does not own (xn_set_interleave) report zero
            XN_CUDA(cudaMemsetAsync(d_steps, 0, n * 4, ctx->stream));
            XN_CUDA(cudaMemsetAsync(d_bytes, 0, n * 8, ctx->stream));
            p.steps_out = d_steps;
            p.bytes_out = d_bytes;
            p.target = d_scratch_target;
            p.target_stride = p.out_w;
            XN_CUDA(xn::launch_traversal(traversal, p, true, ctx->strict, ctx->stream));
            XN_CUDA(xn::launch_stats_totals(d_steps, d_bytes, n, d_tot, ctx->stream));
            if (steps_out) XN_CUDA(cudaMemcpyAsync(steps_out, d_steps, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (bytes_out) XN_CUDA(cudaMemcpyAsync(bytes_out, d_bytes, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
            unsigned long long tot[2] = {0, 0};
            XN_CUDA(cudaMemcpyAsync(tot, d_tot, 16, cudaMemcpyDeviceToHost, ctx->stream));
            XN_CUDA(cudaStreamSynchronize(ctx->stream));
            if (totals_out) {
                totals_out[0] = tot[0];
                totals_out[1] = tot[1];
            }
        } catch (...) {
            cudaFree(d_steps);
            cudaFree(d_bytes);
            cudaFree(d_tot);
            cudaFree(d_scratch_target);
            throw;
        }
        cudaFree(d_steps);
        cudaFree(d_bytes);
        cudaFree(d_tot);
        cudaFree(d_scratch_target);
    });
}

int xn_render_touch_pass(xn_ctx* ctx, const float forward[3], const float up[3], const float translation[3],
                         int use_skip_table, uint64_t counts_out[2]) {
    return guarded([&] {
        check_ctx(ctx);
        if (!counts_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        counts_out[0] = counts_out[1] = 0;
        xn::FrameParams p;
        fill_params(ctx, XN_DDA, forward, up, translation, p);
        if (p.tex_unorm == 0ull)
            throw xn::Error(XN_ERR_INVALID, "the touch pass instruments the texture residency (xn_set_grid_layout)");
        if (!use_skip_table) p.skip_table = nullptr; // every step fetches, as dda.comp:45 does
        DeviceGuard g(ctx->device);
        const uint64_t n = (uint64_t)p.out_w * p.out_h;
        if (n == 0) return;
        const uint64_t voxels = (uint64_t)p.nx * p.ny * p.nz, words = (voxels + 31) / 32;
        uint32_t *d_bits = nullptr, *d_target = nullptr;
        unsigned long long* d_cnt = nullptr;
        try {
            XN_CUDA(cudaMalloc(&d_bits, words * 4));
            XN_CUDA(cudaMalloc(&d_target, n * 4)); // a previous xn_render's image stays untouched
            XN_CUDA(cudaMalloc(&d_cnt, 16));
            XN_CUDA(cudaMemsetAsync(d_bits, 0, words * 4, ctx->stream));
            XN_CUDA(cudaMemsetAsync(d_cnt, 0, 16, ctx->stream));
            p.touch_bits = d_bits;
            p.target = d_target;
            p.target_stride = p.out_w;
            XN_CUDA(xn::launch_traversal(XN_DDA, p, true, ctx->strict, ctx->stream));
            XN_CUDA(xn::launch_touch_count(d_bits, words, d_cnt, ctx->stream));
            unsigned long long h[2] = {0, 0};
            XN_CUDA(cudaMemcpyAsync(h, d_cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
            XN_CUDA(cudaStreamSynchronize(ctx->stream));
            counts_out[0] = h[0];
            counts_out[1] = h[1];
        } catch (...) {
            cudaFree(d_bits);
            cudaFree(d_target);
            cudaFree(d_cnt);
            throw;
        }
        cudaFree(d_bits);
        cudaFree(d_target);
        cudaFree(d_cnt);
    });
}

// ---- multi-device gather ----

int xn_frame_gather(xn_ctx* const* ctxs, int n, uint32_t* host_dst, xn_rect* enclosing_out) {
    return guarded([&] {
        if (!ctxs || n <= 0 || !host_dst) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        for (int i = 0; i < n; ++i) {
            check_ctx(ctxs[i]);
            if (!ctxs[i]->have_target) throw xn::Error(XN_ERR_INVALID, "context without a target");
        }
        xn_rect enc = ctxs[0]->output;
        for (int i = 1; i < n; ++i) enc = xn::rect_union(enc, ctxs[i]->output);
        if (enclosing_out) *enclosing_out = enc;
        const uint64_t px = (uint64_t)enc.w * enc.h;
        if (px == 0) return;
        xn_ctx* root = ctxs[0];
        DeviceGuard g(root->device);
        uint32_t* frame = nullptr;
        XN_CUDA(cudaMalloc(&frame, px * 4));
        try {
            // background 0xFF000000 (BLACK_PIXEL, HeadlessDisplay.cpp:11): memset 0 then alpha
            // via a 2-D memset of the top byte is awkward; fill from the host once instead.
            std::vector<uint32_t> bg;
            bool covered = false;
            for (int i = 0; i < n && !covered; ++i)
                covered = ctxs[i]->output.w == enc.w && ctxs[i]->output.h == enc.h;
            if (!covered) {
                bg.assign(px, 0xFF000000u);
                XN_CUDA(cudaMemcpyAsync(frame, bg.data(), px * 4, cudaMemcpyHostToDevice, root->stream));
                XN_CUDA(cudaStreamSynchronize(root->stream));
            }
            for (int i = 0; i < n; ++i) {
                const xn_ctx* c = ctxs[i];
                if (c->output.w == 0 || c->output.h == 0) continue;
                const uint32_t* src = c->ext_target ? c->ext_target : c->own_target;
                const size_t src_stride = c->ext_target ? c->ext_stride : c->output.w;
                uint32_t* dst = frame + (uint64_t)(c->output.y - enc.y) * enc.w + (uint64_t)(c->output.x - enc.x);
                if (c != root) {
                    // the tile is read on the root's stream: order the read after the work already
                    // enqueued on the tile's own stream (the caller need not have synchronised it)
                    DeviceGuard gc(c->device);
                    XN_CUDA(cudaEventRecord(c->ev_gather, c->stream));
                    XN_CUDA(cudaStreamWaitEvent(root->stream, c->ev_gather, 0));
                }
                // UVA peer copy: goes over NVLink when peer access is possible, staged otherwise
                XN_CUDA(cudaMemcpy2DAsync(dst, (size_t)enc.w * 4, src, src_stride * 4, (size_t)c->output.w * 4,
                                          c->output.h, cudaMemcpyDefault, root->stream));
            }
            XN_CUDA(cudaMemcpyAsync(host_dst, frame, px * 4, cudaMemcpyDeviceToHost, root->stream));
            XN_CUDA(cudaStreamSynchronize(root->stream));
        } catch (...) {
            cudaFree(frame);
            throw;
        }
        cudaFree(frame);
    });
}

int xn_frame_buffer_create(xn_ctx* ctx, uint32_t w, uint32_t h, void** device_ptr_out, uint8_t handle_out[64]) {
    return guarded([&] {
        check_ctx(ctx);
        if (!device_ptr_out || w == 0 || h == 0) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        DeviceGuard g(ctx->device);
        void* p = nullptr;
        const uint64_t px = (uint64_t)w * h;
        XN_CUDA(cudaMalloc(&p, px * 4));
        std::vector<uint32_t> bg(px, 0xFF000000u);
        XN_CUDA(cudaMemcpy(p, bg.data(), px * 4, cudaMemcpyHostToDevice));
        if (handle_out) {
            cudaIpcMemHandle_t hnd;
            cudaError_t e = cudaIpcGetMemHandle(&hnd, p);
            if (e != cudaSuccess) {
                cudaFree(p);
                throw CudaError{e, "cudaIpcGetMemHandle"};
            }
            std::memcpy(handle_out, &hnd, 64);
        }
        *device_ptr_out = p;
    });
}

int xn_frame_buffer_open(xn_ctx* ctx, const uint8_t handle[64], void** device_ptr_out) {
    return guarded([&] {
        check_ctx(ctx);
        if (!handle || !device_ptr_out) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        DeviceGuard g(ctx->device);
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, handle, 64);
        void* p = nullptr;
        XN_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_opened.push_back(p);
        *device_ptr_out = p;
    });
}

int xn_frame_buffer_close(xn_ctx* ctx, void* device_ptr) {
    return guarded([&] {
        check_ctx(ctx);
        if (!device_ptr) return;
        DeviceGuard g(ctx->device);
        auto it = std::find(ctx->ipc_opened.begin(), ctx->ipc_opened.end(), device_ptr);
        if (it != ctx->ipc_opened.end()) {
            ctx->ipc_opened.erase(it);
            XN_CUDA(cudaIpcCloseMemHandle(device_ptr));
        } else {
            XN_CUDA(cudaFree(device_ptr));
        }
    });
}

int xn_frame_buffer_read(xn_ctx* ctx, const void* device_ptr, uint32_t w, uint32_t h, uint32_t* host_dst) {
    return guarded([&] {
        check_ctx(ctx);
        if (!device_ptr || !host_dst) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        DeviceGuard g(ctx->device);
        XN_CUDA(cudaMemcpyAsync(host_dst, device_ptr, (uint64_t)w * h * 4, cudaMemcpyDeviceToHost, ctx->stream));
        XN_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int xn_frame_buffer_read_async(xn_ctx* ctx, const void* device_ptr, uint32_t w, uint32_t h, uint32_t* host_dst) {
    return guarded([&] {
        check_ctx(ctx);
        if (!device_ptr || !host_dst) throw xn::Error(XN_ERR_INVALID, "bad arguments");
        DeviceGuard g(ctx->device);
        // ordered after everything already enqueued on the context's stream, but executed on
        // the copy stream so that later traversal launches overlap it
        XN_CUDA(cudaEventRecord(ctx->ev_rendered[0], ctx->stream));
        XN_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[0], 0));
        XN_CUDA(cudaMemcpyAsync(host_dst, device_ptr, (uint64_t)w * h * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
        ctx->frame_read_in_flight = true;
    });
}

int xn_copy_sync(xn_ctx* ctx) {
    return guarded([&] {
        check_ctx(ctx);
        DeviceGuard g(ctx->device);
        XN_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        for (bool& f : ctx->copy_in_flight) f = false;
        ctx->frame_read_in_flight = false;
    });
}

// ---- host formats ----

int xn_tiff_info(const char* path, uint64_t dims_out[3]) {
    return guarded([&] {
        if (!path || !dims_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        const auto info = xn::tiff_info(path);
        dims_out[0] = info.nx;
        dims_out[1] = info.ny;
        dims_out[2] = info.nz;
    });
}
int xn_tiff_stream_info(const char* path, int* streamable_out, uint32_t format_out[5], uint64_t* runs_out) {
    return guarded([&] {
        if (!path || !streamable_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        const xn::TiffPlan plan = xn::tiff_plan(path);
        *streamable_out = plan.streamable ? 1 : 0;
        if (format_out) {
            format_out[0] = plan.format.samples;
            format_out[1] = plan.format.photometric;
            format_out[2] = plan.format.has_alpha;
            format_out[3] = plan.format.unassociated;
            format_out[4] = plan.format.flip;
        }
        if (runs_out) {
            *runs_out = 0;
            for (const auto& sl : plan.slices) *runs_out += sl.size();
        }
    });
}
int xn_tiff_read(const char* path, uint8_t* rgba_out, uint64_t cap_bytes) {
    return guarded([&] {
        if (!path || !rgba_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        xn::tiff_read(path, rgba_out, cap_bytes);
    });
}
int xn_tiff_write(const char* path, const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, int bigtiff) {
    return guarded([&] {
        if (!path || !rgba) throw xn::Error(XN_ERR_INVALID, "null argument");
        xn::tiff_write(path, rgba, nx, ny, nz, bigtiff != 0);
    });
}
int xn_svo_info(const char* path, uint64_t* side_out, uint64_t* count_out) {
    return guarded([&] {
        if (!path) throw xn::Error(XN_ERR_INVALID, "null argument");
        uint64_t side, count;
        xn::svo_info(path, side, count);
        if (side_out) *side_out = side;
        if (count_out) *count_out = count;
    });
}
int xn_svo_read(const char* path, xn_node* nodes_out, uint64_t cap_nodes) {
    return guarded([&] {
        if (!path || !nodes_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        xn::svo_read(path, nodes_out, cap_nodes);
    });
}
int xn_svo_write(const char* path, const xn_node* nodes, uint64_t count, uint64_t side) {
    return guarded([&] {
        if (!path || !nodes) throw xn::Error(XN_ERR_INVALID, "null argument");
        xn::save_svo(path, nodes, count, side);
    });
}

int xn_build_octree(const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, int heuristic, double param, int type,
                    xn_node** nodes_out, uint64_t* count_out, uint64_t* side_out, xn_build_stats* stats_out) {
    return guarded([&] {
        if (!rgba || !nodes_out || !count_out || !side_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        if (heuristic < 0 || heuristic > 1 || type < 0 || type > 2) throw xn::Error(XN_ERR_INVALID, "bad option");
        xn::Octree t = xn::build_octree(rgba, nx, ny, nz, (xn::Heuristic)heuristic, param, (xn::OctreeType)type, stats_out);
        xn_node* out = (xn_node*)std::malloc(std::max<size_t>(t.nodes.size(), 1) * sizeof(xn_node));
        if (!out) throw std::bad_alloc();
        std::memcpy(out, t.nodes.data(), t.nodes.size() * sizeof(xn_node));
        *nodes_out = out;
        *count_out = t.nodes.size();
        *side_out = t.side;
    });
}

void xn_free(void* p) { std::free(p); }

int xn_headless_config_parse(const char* text, xn_headless_device* out, int cap, int* count_out) {
    return guarded([&] {
        if (!text || !count_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        const auto devs = xn::parse_headless_config(text);
        *count_out = (int)devs.size();
        if (out)
            for (int i = 0; i < cap && i < (int)devs.size(); ++i) out[i] = devs[i];
    });
}

int xn_camera_script_parse(const char* text, float* frames_out, int cap_frames, int* count_out) {
    return guarded([&] {
        if (!text || !count_out) throw xn::Error(XN_ERR_INVALID, "null argument");
        const auto frames = xn::parse_camera_script(text);
        *count_out = (int)frames.size();
        if (frames_out)
            for (int i = 0; i < cap_frames && i < (int)frames.size(); ++i) {
                std::memcpy(frames_out + 9 * i, frames[i].forward, 12);
                std::memcpy(frames_out + 9 * i + 3, frames[i].up, 12);
                std::memcpy(frames_out + 9 * i + 6, frames[i].translation, 12);
            }
    });
}

int xn_stats_write(const char* path, const xn_render_stats* frames, uint64_t n_frames, double wall_seconds) {
    return guarded([&] {
        if (!path || (!frames && n_frames)) throw xn::Error(XN_ERR_INVALID, "null argument");
        xn::stats_write(path, frames, n_frames, wall_seconds);
    });
}

int xn_png_write(const char* path, const uint32_t* rgba, uint32_t w, uint32_t h) {
    return guarded([&] {
        if (!path || !rgba) throw xn::Error(XN_ERR_INVALID, "null argument");
        xn::png_write(path, rgba, w, h);
    });
}

} // extern "C"
