// xn_synth_host.cpp -- host generator of the synthetic volumes (same integer code as the
// device generator, xn_synth.h), rows spread over hardware threads.
//
// The gas volume costs 48 lattice hashes and 42 integer lerps per voxel when evaluated voxel by
// voxel (two ridged fields x three octaves of trilinear value noise).  Along an x row the lattice
// cell of an octave changes only every 2^shift voxels, so the row generator keeps the eight
// corner values of each of the six noise instances and re-hashes them at cell boundaries; the
// lerps run in the order of synth_value_noise (x, then y, then z), so the voxels are bit-identical
// (tests/test_host_formats.py compares the two paths; XN_SYNTH_PLAIN=1 forces the plain one).
#include <cstdlib>
#include <thread>
#include <vector>

#include "../xn_synth.h"
#include "xn_host.hpp"

namespace xn {
namespace {

struct NoiseRow {
    int shift;
    uint32_t seed, j, k, fy, fz, mask;
    int32_t c[8];
    uint32_t cell; // lattice cell the corners belong to
    void begin(uint32_t y, uint32_t z, int shift_, uint32_t seed_) {
        shift = shift_;
        seed = seed_;
        mask = (1u << shift) - 1u;
        j = y >> shift;
        k = z >> shift;
        fy = (y & mask) << (16 - shift);
        fz = (z & mask) << (16 - shift);
        cell = 0xFFFFFFFFu;
    }
    uint32_t at(uint32_t x) {
        const uint32_t i = x >> shift;
        if (i != cell) {
            cell = i;
            for (int d = 0; d < 8; ++d)
                c[d] = (int32_t)(synth_hash(i + ((d >> 2) & 1), j + ((d >> 1) & 1), k + (d & 1), seed) & 0xFFFFu);
        }
        const uint32_t fx = (x & mask) << (16 - shift);
        const int32_t x00 = synth_lerp(c[0], c[4], fx), x01 = synth_lerp(c[1], c[5], fx);
        const int32_t x10 = synth_lerp(c[2], c[6], fx), x11 = synth_lerp(c[3], c[7], fx);
        const int32_t y0 = synth_lerp(x00, x10, fy), y1 = synth_lerp(x01, x11, fy);
        return (uint32_t)synth_lerp(y0, y1, fz);
    }
};

uint32_t ridge_of(uint32_t o0, uint32_t o1, uint32_t o2) { // synth_ridged
    const int32_t n = (int32_t)((4u * o0 + 2u * o1 + o2) / 7u);
    int32_t d = n - 32768;
    d = d < 0 ? -d : d;
    const int32_t r = 65535 - 4 * d;
    return (uint32_t)(r < 0 ? 0 : r);
}

// one x row of the gas volume; requires s0 - 2 >= 1 (every octave on a lattice coarser than a voxel)
void tng_row(const SynthSpec& s, uint32_t y, uint32_t z, int s0, uint32_t* out) {
    NoiseRow n[6];
    for (int f = 0; f < 2; ++f)
        for (int o = 0; o < 3; ++o) n[f * 3 + o].begin(y, z, s0 - o, s.seed + (f ? 101u : 0u) + (uint32_t)o);
    uint32_t maxdim = s.nx > s.ny ? s.nx : s.ny;
    maxdim = maxdim > s.nz ? maxdim : s.nz;
    const uint32_t lo = maxdim >= 2048u ? 58410u : 54150u, hi = 61960u;
    const uint32_t floor_colour = synth_magma(0);
    for (uint32_t x = 0; x < s.nx; ++x) {
        const uint32_t r1 = ridge_of(n[0].at(x), n[1].at(x), n[2].at(x));
        const uint32_t r2 = ridge_of(n[3].at(x), n[4].at(x), n[5].at(x));
        const uint32_t r = r1 < r2 ? r1 : r2;
        if (r <= lo) {
            out[x] = floor_colour;
        } else {
            const uint32_t t = ((r - lo) * 255u) / (hi - lo);
            out[x] = synth_magma(t > 255u ? 255u : t);
        }
    }
}

} // namespace

void synth_grid_host(int kind, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t seed, uint8_t* rgba_out) {
    if (kind < 0 || kind > 1 || nx == 0 || ny == 0 || nz == 0 || nx > 0xFFFFu || ny > 0xFFFFu || nz > 0xFFFFu)
        throw Error(XN_ERR_INVALID, "synth_grid_host: bad volume specification");
    const SynthSpec spec{(uint32_t)kind, (uint32_t)nx, (uint32_t)ny, (uint32_t)nz, seed};
    uint32_t* out = reinterpret_cast<uint32_t*>(rgba_out);
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > nz) nt = (unsigned)nz;
    uint32_t maxdim = spec.nx > spec.ny ? spec.nx : spec.ny;
    maxdim = maxdim > spec.nz ? maxdim : spec.nz;
    const int s0 = synth_ilog2(maxdim) - 3;
    const char* plain = std::getenv("XN_SYNTH_PLAIN");
    const bool rows = kind == 1 && s0 - 2 >= 1 && !(plain && plain[0] == '1');
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([=] {
            for (uint64_t z = t; z < nz; z += nt)
                for (uint64_t y = 0; y < ny; ++y) {
                    uint32_t* row = out + (y * nx + z * nx * ny);
                    if (rows) {
                        tng_row(spec, (uint32_t)y, (uint32_t)z, s0, row);
                    } else {
                        for (uint64_t x = 0; x < nx; ++x) row[x] = synth_voxel(spec, (uint32_t)x, (uint32_t)y, (uint32_t)z);
                    }
                }
        });
    for (auto& th : pool) th.join();
}

} // namespace xn
