// xn_host.hpp -- host-side C++ mirror of the reference's data model for the traversal path
// (grid / octree containers, file formats, headless config, camera script, stats).
// Everything here is plain C++17 with no CUDA dependency; the C ABI (xn_capi.cu) and the
// `xenodon` CLI (xn_cli.cpp) are thin layers over it.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "xenodon_b200.h"

namespace xn {

// Error carrying an xn_status; mirrors the reference's `Error : std::runtime_error`
// (src/core/Error.h:8-13) with the same message texts where the reference defines them.
struct Error : std::runtime_error {
    int status;
    Error(int status_, const std::string& msg) : std::runtime_error(msg), status(status_) {}
};

// ---- Grid (src/model/Grid.h:13-65): RGBA8, index x + y*nx + z*nx*ny ----
struct Grid {
    uint64_t nx = 0, ny = 0, nz = 0;
    std::vector<uint8_t> rgba; // 4 * nx * ny * nz
    uint64_t voxels() const { return nx * ny * nz; }
};

struct TiffInfo {
    uint64_t nx, ny, nz;
};
TiffInfo tiff_info(const std::string& path);
// Grid::load_tiff (src/model/Grid.cpp:27-79) without libtiff; `out` must hold 4*nx*ny*nz bytes
void tiff_read(const std::string& path, uint8_t* out, uint64_t cap_bytes);
Grid load_tiff(const std::string& path);
// Streaming ingest (volume upload pipeline): where the raw sample bytes of each z slice lie in
// the file and how they decode, so that slices can be read straight into page-locked staging
// memory and decoded (sample expansion, alpha pre-multiplication, row flip) on the device.
struct TiffSliceFormat { // identical for every slice of a plan that is `streamable`
    uint32_t samples;        // 1..4 bytes per pixel
    uint32_t photometric;    // 0 = white-is-zero, 1 = black-is-zero, 2 = RGB
    uint32_t has_alpha, unassociated;
    uint32_t flip;           // Orientation tag: bit 0 file row r is grid row H - 1 - r, bit 1 columns reversed too
};
struct TiffRun { // `bytes` contiguous file bytes at `offset`: whole rows in file order
    uint64_t offset, bytes;
};
struct TiffPlan {
    TiffInfo info{};
    bool streamable = false; // uncompressed chunky 8-bit strips with one format throughout
    TiffSliceFormat format{};
    std::vector<std::vector<TiffRun>> slices; // per z: runs covering nx*ny*samples bytes
};
TiffPlan tiff_plan(const std::string& path);
// host decode of one slice (any supported layout) into out[nx*ny*4], the fallback of the pipeline
void tiff_read_slice(const std::string& path, uint64_t z, uint8_t* out);
void tiff_write(const std::string& path, const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, bool bigtiff);

// ---- Octree (src/model/Octree.h, Octree.cpp:50-114) ----
struct Octree {
    uint64_t side = 0;
    std::vector<xn_node> nodes;
};
void svo_info(const std::string& path, uint64_t& side, uint64_t& count);
void svo_read(const std::string& path, xn_node* out, uint64_t cap_nodes);
Octree load_svo(const std::string& path);
void save_svo(const std::string& path, const xn_node* nodes, uint64_t count, uint64_t side);

// ---- build_octree (src/model/OctreeConstruction.h:226-237) ----
enum class OctreeType { Sparse = 0, Dag = 1, Rope = 2 };
enum class Heuristic { ChanDiff = 0, StdDev = 1 };
Octree build_octree(const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, Heuristic h, double param,
                    OctreeType type, xn_build_stats* stats);
void generate_ropes(Octree& tree); // src/model/Octree.cpp:181-201

// ---- headless.conf (src/backend/headless/HeadlessConfig.cpp:5-28, src/core/Config.h, Parser.cpp) ----
std::vector<xn_headless_device> parse_headless_config(const std::string& text);
// RenderContext::calculate_display_rect / rect_union (src/utility/rect_union.h:11-26)
xn_rect rect_union(const xn_rect& a, const xn_rect& b);

// ---- camera script (src/camera/ScriptCameraController.cpp:4-41) ----
struct CameraFrame {
    float forward[3], up[3], translation[3];
};
std::vector<CameraFrame> parse_camera_script(const std::string& text);

// ---- stats (src/render/RenderStats.cpp:13-24, 85-159) ----
void stats_combine(xn_render_stats& into, const xn_render_stats& other);
double stats_mrays_per_s(const xn_render_stats& s);
std::string stats_format(const std::vector<xn_render_stats>& frames, double wall_seconds);
void stats_write(const std::string& path, const xn_render_stats* frames, uint64_t n, double wall_seconds);

// ---- PNG (replaces lodepng::encode in HeadlessDisplay::save, HeadlessDisplay.cpp:78-83) ----
void png_write(const std::string& path, const uint32_t* rgba, uint32_t w, uint32_t h);

// ---- synthetic volumes (xn_synth.h) on the host ----
void synth_grid_host(int kind, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t seed, uint8_t* rgba_out);

// `{}`-style formatting of doubles as fmt prints them (shortest round-trip)
std::string fmt_double(double v);

} // namespace xn
