// adapter_check.cpp -- compiles integration/CudaMultiplexRenderer.h against the reference's own
// headers and exercises it: a 16^3 grid, one device block, one camera.  Prints "rendered <n> rays
// <checksum>" when a CUDA device is present, "error: <library message>" when not (exit 3).
// Built and run by tests/test_integration_adapter.py; never part of the product.
#include <cstdio>
#include <memory>

#include "CudaMultiplexRenderer.h"

int main() {
    Grid grid(Vec3Sz{16, 16, 16});
    for (size_t z = 0; z < 16; ++z)
        for (size_t y = 0; y < 16; ++y)
            for (size_t x = 0; x < 16; ++x)
                grid.set({x, y, z}, Pixel{uint8_t(16 * x), uint8_t(16 * y), uint8_t(16 * z), 255});
    HeadlessConfig cfg;
    HeadlessConfig::Device dev;
    dev.vulkan_index = 0;
    dev.region.offset.x = 0, dev.region.offset.y = 0;
    dev.region.extent.width = 64, dev.region.extent.height = 36;
    cfg.gpus.push_back(dev);
    try {
        CudaMultiplexRenderer r(cfg, "dda", &grid, nullptr, Vec3F{1.f, 1.f, 1.f}, 2.0f);
        Camera cam;
        cam.forward = Vec3F{0.f, 0.f, 1.f};
        cam.up = Vec3F{0.f, 1.f, 0.f};
        cam.translation = Vec3F{0.5f, 0.5f, -1.5f};
        r.render(cam);
        const auto image = r.frame();
        unsigned long long sum = 0;
        for (const Pixel& p : image) sum += p.r + p.g + p.b;
        std::printf("rendered %zu rays %llu\n", r.stats().total_rays, sum);
        return sum != 0 && r.stats().total_rays == 64 * 36 ? 0 : 4;
    } catch (const std::exception& e) {
        std::printf("error: %s\n", xn_last_error());
        return 3;
    }
}
