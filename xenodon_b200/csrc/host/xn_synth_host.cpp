// xn_synth_host.cpp -- host generator of the synthetic volumes (same integer code as the
// device generator, xn_synth.h), rows spread over hardware threads.
#include <thread>

#include "../xn_synth.h"
#include "xn_host.hpp"

namespace xn {

void synth_grid_host(int kind, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t seed, uint8_t* rgba_out) {
    if (kind < 0 || kind > 1 || nx == 0 || ny == 0 || nz == 0 || nx > 0xFFFFu || ny > 0xFFFFu || nz > 0xFFFFu)
        throw Error(XN_ERR_INVALID, "synth_grid_host: bad volume specification");
    const SynthSpec spec{(uint32_t)kind, (uint32_t)nx, (uint32_t)ny, (uint32_t)nz, seed};
    uint32_t* out = reinterpret_cast<uint32_t*>(rgba_out);
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > nz) nt = (unsigned)nz;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([=] {
            for (uint64_t z = t; z < nz; z += nt)
                for (uint64_t y = 0; y < ny; ++y)
                    for (uint64_t x = 0; x < nx; ++x)
                        out[x + y * nx + z * nx * ny] = synth_voxel(spec, (uint32_t)x, (uint32_t)y, (uint32_t)z);
        });
    for (auto& th : pool) th.join();
}

} // namespace xn
