// ref_model.cpp -- C entry points around the reference's OWN model code
// (src/model/Grid.cpp, Octree.cpp, OctreeConstruction.h compiled where they lie under
// /root/reference).  Used to pin our TIFF reader, .svo reader/writer and `convert`
// restatements byte-for-byte.  TEST INFRASTRUCTURE ONLY.
#include <cstring>
#include <exception>
#include "model/Grid.h"
#include "model/Octree.h"
#include "model/OctreeConstruction.h"

struct xnref_build_stats {
    uint64_t total_leaves, unique_leaves, total_nodes, depth;
};

static Octree build(const Grid& grid, int heuristic, double param, int type, ConstructionStats& st) {
    auto t = type == 1 ? Octree::Type::Dag : type == 2 ? Octree::Type::Rope : Octree::Type::Sparse;
    if (heuristic == 1) return build_octree(grid, st, StdDevHeuristic{param}, t);
    return build_octree(grid, st, ChannelDiffHeuristic{static_cast<uint8_t>(param)}, t);
}

// what `xenodon convert` does (src/convert.cpp:56-121): load_tiff -> build_octree -> save_svo
extern "C" int xnref_convert(const char* tif, const char* svo, int heuristic, double param, int type,
                             xnref_build_stats* out) {
    try {
        Grid grid = Grid::load_tiff(tif);
        ConstructionStats st;
        Octree oct = build(grid, heuristic, param, type, st);
        oct.save_svo(svo);
        if (out) *out = {st.total_leaves, st.unique_leaves, st.total_nodes, st.depth};
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// same, from an in-memory RGBA8 grid (x fastest) to an .svo file
extern "C" int xnref_convert_mem(const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, const char* svo,
                                 int heuristic, double param, int type, xnref_build_stats* out) {
    try {
        Grid grid(Vec3Sz{nx, ny, nz});
        for (uint64_t z = 0; z < nz; ++z)
            for (uint64_t y = 0; y < ny; ++y)
                for (uint64_t x = 0; x < nx; ++x) {
                    const uint8_t* p = rgba + 4 * (x + y * nx + z * nx * ny);
                    grid.set({x, y, z}, Pixel{p[0], p[1], p[2], p[3]});
                }
        ConstructionStats st;
        Octree oct = build(grid, heuristic, param, type, st);
        oct.save_svo(svo);
        if (out) *out = {st.total_leaves, st.unique_leaves, st.total_nodes, st.depth};
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// Grid::load_tiff; dims_out = {w, h, d}; rgba_out (nullable) receives w*h*d*4 bytes
extern "C" int xnref_load_tiff(const char* tif, uint64_t dims_out[3], uint8_t* rgba_out, uint64_t cap) {
    try {
        Grid grid = Grid::load_tiff(tif);
        auto d = grid.dimensions();
        dims_out[0] = d.x; dims_out[1] = d.y; dims_out[2] = d.z;
        if (rgba_out) {
            if (cap < grid.size() * 4) return -2;
            std::memcpy(rgba_out, grid.pixels().data(), grid.size() * 4);
        }
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// Octree::load_svo followed by Octree::save_svo (round trip through the reference's reader/writer)
extern "C" int xnref_svo_roundtrip(const char* in, const char* out, uint64_t* side, uint64_t* count) {
    try {
        Octree oct = Octree::load_svo(in);
        if (side) *side = oct.side();
        if (count) *count = oct.data().size();
        if (out) oct.save_svo(out);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
