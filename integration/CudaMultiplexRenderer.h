// CudaMultiplexRenderer.h -- the reference-side adapter INTEGRATION.md describes: one class that a
// maintainer drops into the reference tree (src/render/) in place of MultiplexRenderer + the
// headless display, forwarding the traversal path to libxenodon_b200.so through its C ABI.
//
// It is written against the reference's OWN headers (core/Error.h, camera/Camera.h,
// render/RenderStats.h, backend/headless/HeadlessConfig.h, model/Grid.h, model/Octree.h) and
// include/xenodon_b200.h, nothing else.  tests/test_integration_adapter.py compiles it against
// those headers where the reference checkout is present (the Vulkan SDK and fmt are stood in for by
// integration/shim), checks that every xn_* symbol it binds is exported by the library, and runs
// it: with a CUDA device it renders a frame, without one it fails loudly with the library's error.
//
// Replaces, in the reference:
//   create_headless_display + MultiplexRenderer ctor   src/main_loop.cpp:203, src/render/MultiplexRenderer.cpp:5-19
//   MultiplexRenderer::render + Display::swap_buffers  src/render/MultiplexRenderer.cpp:21-31,
//                                                      src/backend/headless/HeadlessDisplay.cpp:39-54
//   MultiplexRenderer::stats                           src/render/MultiplexRenderer.cpp:33-41
//   HeadlessDisplay::save's tile composite             src/backend/headless/HeadlessDisplay.cpp:59-76
#pragma once
#include <algorithm>
#include <string>
#include <string_view>
#include <vector>

#include <xenodon_b200.h>

#include "backend/headless/HeadlessConfig.h"
#include "camera/Camera.h"
#include "core/Error.h"
#include "model/Grid.h"
#include "model/Octree.h"
#include "model/Pixel.h"
#include "render/RenderStats.h"

class CudaMultiplexRenderer {
    struct Dev {
        xn_ctx* ctx;
        xn_rect region;
    };
    std::vector<Dev> devs;
    xn_rect display{0, 0, 0, 0};
    int traversal;
    RenderStats last;

    static void check(int rc) {
        if (rc != XN_OK) throw Error("{}", xn_last_error());
    }

    // rect_union, src/utility/rect_union.h:11-26
    static xn_rect unite(const xn_rect& a, const xn_rect& b) {
        const int32_t x = std::min(a.x, b.x), y = std::min(a.y, b.y);
        const int64_t r = std::max<int64_t>(int64_t(a.x) + a.w, int64_t(b.x) + b.w);
        const int64_t t = std::max<int64_t>(int64_t(a.y) + a.h, int64_t(b.y) + b.h);
        return {x, y, uint32_t(r - x), uint32_t(t - y)};
    }

public:
    // One context per `device { vkindex offset extent }` block of the headless config (vkindex =
    // CUDA ordinal), volume replicated on each: DdaRaytraceAlgorithm / SvoRaytraceAlgorithm
    // ::upload_resources (src/render/DdaRaytraceAlgorithm.cpp:15-97, SvoRaytraceAlgorithm.cpp:12-48).
    CudaMultiplexRenderer(const HeadlessConfig& cfg, std::string_view shader, const Grid* grid, const Octree* octree,
                          Vec3F voxel_ratio, float emission) {
        traversal = xn_traversal_from_name(std::string(shader).c_str());
        if (traversal < 0) throw Error("{}", xn_last_error());
        if (cfg.gpus.empty()) throw Error("{}", "headless config names no device");
        bool first = true;
        for (const auto& gpu : cfg.gpus) {
            const xn_rect r{gpu.region.offset.x, gpu.region.offset.y, gpu.region.extent.width, gpu.region.extent.height};
            Dev d{nullptr, r};
            check(xn_ctx_create(int(gpu.vulkan_index), &d.ctx));
            devs.push_back(d);
            display = first ? r : unite(display, r);
            first = false;
        }
        for (auto& d : devs) {
            uint32_t dim[3];
            if (grid) {
                const auto s = grid->dimensions();
                check(xn_upload_grid(d.ctx, reinterpret_cast<const uint8_t*>(grid->pixels().data()), s.x, s.y, s.z));
                dim[0] = uint32_t(s.x), dim[1] = uint32_t(s.y), dim[2] = uint32_t(s.z);
            } else {
                static_assert(sizeof(Octree::Node) == sizeof(xn_node), "the .svo node is the C ABI's xn_node");
                check(xn_upload_svo(d.ctx, reinterpret_cast<const xn_node*>(octree->data().data()), octree->data().size(),
                                    octree->side()));
                dim[0] = dim[1] = dim[2] = uint32_t(octree->side());
            }
            const float ratio[3] = {voxel_ratio.x, voxel_ratio.y, voxel_ratio.z};
            check(xn_set_target(d.ctx, &d.region, &display));
            check(xn_set_params(d.ctx, ratio, dim, emission));
        }
    }
    CudaMultiplexRenderer(const CudaMultiplexRenderer&) = delete;
    CudaMultiplexRenderer& operator=(const CudaMultiplexRenderer&) = delete;
    ~CudaMultiplexRenderer() {
        for (auto& d : devs) xn_ctx_destroy(d.ctx);
    }

    // launches on every device first, then waits for all of them (swap_buffers' fence wait) and
    // folds the per-device kernel times the way RenderStats::combine does (RenderStats.cpp:13-20)
    void render(const Camera& cam) {
        const float f[3] = {cam.forward.x, cam.forward.y, cam.forward.z};
        const float u[3] = {cam.up.x, cam.up.y, cam.up.z};
        const float t[3] = {cam.translation.x, cam.translation.y, cam.translation.z};
        for (auto& d : devs) check(xn_render(d.ctx, traversal, f, u, t));
        last = RenderStats{};
        for (auto& d : devs) {
            double ms = 0;
            check(xn_sync(d.ctx, &ms));
            last.total_rays += size_t(d.region.w) * d.region.h;
            last.outputs += 1;
            last.total_render_time += ms;
            last.max_render_time = std::max(last.max_render_time, ms);
            last.min_render_time = std::min(last.min_render_time, ms);
        }
    }
    const RenderStats& stats() const { return last; }
    xn_rect enclosing() const { return display; }

    // the composited frame HeadlessDisplay::save hands to lodepng
    std::vector<Pixel> frame() {
        std::vector<xn_ctx*> c;
        for (auto& d : devs) c.push_back(d.ctx);
        std::vector<Pixel> image(size_t(display.w) * display.h);
        xn_rect enc{};
        check(xn_frame_gather(c.data(), int(c.size()), reinterpret_cast<uint32_t*>(image.data()), &enc));
        return image;
    }
};
