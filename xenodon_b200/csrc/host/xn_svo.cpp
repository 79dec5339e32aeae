// xn_svo.cpp -- .svo reader/writer and `xenodon convert`'s octree construction.
//
// File format (reference src/model/Octree.cpp:50-114, src/utility/serialization.h:10-38):
//   "XNDN-SVO", u64 LE side, u64 LE node count, count x { 8 x u32 children, u32 colour,
//   u32 is_leaf_depth }, all little endian; file size must be exactly 24 + 40 * count.
//
// Construction (reference src/model/OctreeConstruction.h:124-237, Grid.cpp:81-214): the node
// ARRAY must come out byte-identical to the reference's recursive top-down builder (children
// inserted post-order in x-major order, array reversed, child indices mirrored).  The
// reference rescans every region at every level (O(N * depth)); here the --chan-diff path
// builds bottom-up above a small block size: children are emitted first and rolled back when
// the parent turns out to be a leaf, so every voxel is read O(1) times.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <unordered_map>

#include "xn_host.hpp"

namespace xn {
namespace {

constexpr uint32_t LEAF = 0x80000000u;
const char SVO_FMT_ID[8] = {'X', 'N', 'D', 'N', '-', 'S', 'V', 'O'};

struct File {
    FILE* f;
    File(const std::string& p, const char* m) : f(std::fopen(p.c_str(), m)) {}
    ~File() {
        if (f) std::fclose(f);
    }
};

uint64_t le64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}

void read_header(FILE* f, uint64_t& side, uint64_t& count) {
    uint8_t hdr[24];
    if (std::fread(hdr, 1, 24, f) != 24 || std::memcmp(hdr, SVO_FMT_ID, 8) != 0)
        throw Error(XN_ERR_FORMAT, "Invalid format id");
    side = le64(hdr + 8);
    count = le64(hdr + 16);
    if (fseeko(f, 0, SEEK_END) != 0) throw Error(XN_ERR_IO, "Failed to tell");
    const off_t end = ftello(f);
    if (end < 0) throw Error(XN_ERR_IO, "Failed to tell");
    if ((uint64_t)end - 24 != count * sizeof(xn_node) || count > (uint64_t)end)
        throw Error(XN_ERR_FORMAT, "File size does not match number of nodes");
    fseeko(f, 24, SEEK_SET);
}

} // namespace

void svo_info(const std::string& path, uint64_t& side, uint64_t& count) {
    File file(path, "rb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    read_header(file.f, side, count);
}

void svo_read(const std::string& path, xn_node* out, uint64_t cap_nodes) {
    File file(path, "rb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    uint64_t side, count;
    read_header(file.f, side, count);
    if (cap_nodes < count) throw Error(XN_ERR_INVALID, "svo_read: output buffer too small");
    // x86-64 / little-endian host (as the reference requires, meson.build): the on-disk node
    // is the in-memory node, so the array is read in bulk instead of field by field.
    static_assert(sizeof(xn_node) == 40, "xn_node must be 40 bytes");
    if (std::fread(out, sizeof(xn_node), count, file.f) != count) throw Error(XN_ERR_IO, "svo_read: short read");
}

Octree load_svo(const std::string& path) {
    Octree t;
    uint64_t count;
    svo_info(path, t.side, count);
    t.nodes.resize(count);
    svo_read(path, t.nodes.data(), count);
    return t;
}

void save_svo(const std::string& path, const xn_node* nodes, uint64_t count, uint64_t side) {
    File file(path, "wb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    uint8_t hdr[24];
    std::memcpy(hdr, SVO_FMT_ID, 8);
    for (int i = 0; i < 8; ++i) {
        hdr[8 + i] = (uint8_t)(side >> (8 * i));
        hdr[16 + i] = (uint8_t)(count >> (8 * i));
    }
    if (std::fwrite(hdr, 1, 24, file.f) != 24 || std::fwrite(nodes, sizeof(xn_node), count, file.f) != count ||
        std::fflush(file.f) != 0)
        throw Error(XN_ERR_IO, "save_svo: short write");
}

// ---------------------------------------------------------------------------------
// construction
// ---------------------------------------------------------------------------------
namespace {

struct NodeHash {
    size_t operator()(const xn_node& n) const {
        uint64_t h = 1469598103934665603ull;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&n);
        for (int i = 0; i < 10; ++i) {
            h ^= w[i];
            h *= 1099511628211ull;
        }
        return (size_t)h;
    }
};
struct NodeEq {
    bool operator()(const xn_node& a, const xn_node& b) const { return std::memcmp(&a, &b, sizeof(xn_node)) == 0; }
};

// per-region summary: what Grid::vol_scan (Grid.cpp:81-137) derives its result from
struct Summary {
    uint64_t sum[4] = {0, 0, 0, 0};
    uint64_t n = 0;
    uint8_t mn[4] = {255, 255, 255, 255}, mx[4] = {0, 0, 0, 0};
    void add(const Summary& o) {
        for (int c = 0; c < 4; ++c) {
            sum[c] += o.sum[c];
            mn[c] = std::min(mn[c], o.mn[c]);
            mx[c] = std::max(mx[c], o.mx[c]);
        }
        n += o.n;
    }
    uint32_t avg_color() const {
        if (n == 0) return 0;
        uint32_t c = 0;
        for (int k = 0; k < 4; ++k) c |= (uint32_t)(uint8_t)(sum[k] / n) << (8 * k);
        return c;
    }
    uint8_t max_diff() const {
        if (n == 0) return 0;
        uint8_t d = 0;
        for (int c = 0; c < 4; ++c) d = std::max<uint8_t>(d, (uint8_t)(mx[c] - mn[c]));
        return d;
    }
};

struct Builder {
    const uint8_t* grid;
    uint64_t nx, ny, nz;
    Heuristic heuristic;
    double param;
    bool dag;
    std::vector<xn_node> nodes;
    std::unordered_map<xn_node, uint32_t, NodeHash, NodeEq> cache;
    xn_build_stats stats{0, 0, 0, 0};

    // direct scan of [o, o+extent) clipped to the grid
    Summary scan(const uint64_t o[3], uint64_t extent) const {
        Summary s;
        const uint64_t x0 = std::min(nx, o[0]), y0 = std::min(ny, o[1]), z0 = std::min(nz, o[2]);
        const uint64_t x1 = std::min(nx, o[0] + extent), y1 = std::min(ny, o[1] + extent),
                       z1 = std::min(nz, o[2] + extent);
        for (uint64_t z = z0; z < z1; ++z)
            for (uint64_t y = y0; y < y1; ++y) {
                const uint8_t* row = grid + 4 * (y * nx + z * nx * ny);
                for (uint64_t x = x0; x < x1; ++x)
                    for (int c = 0; c < 4; ++c) {
                        const uint8_t v = row[4 * x + c];
                        s.sum[c] += v;
                        s.mn[c] = std::min(s.mn[c], v);
                        s.mx[c] = std::max(s.mx[c], v);
                    }
            }
        s.n = (x1 - x0) * (y1 - y0) * (z1 - z0);
        return s;
    }

    // Grid::stddev_scan (Grid.cpp:139-214): same two-pass double arithmetic, same order
    double stddev_of(const uint64_t o[3], uint64_t extent, const Summary& s) const {
        if (s.n == 0) return 0.0;
        const uint64_t x0 = std::min(nx, o[0]), y0 = std::min(ny, o[1]), z0 = std::min(nz, o[2]);
        const uint64_t x1 = std::min(nx, o[0] + extent), y1 = std::min(ny, o[1] + extent),
                       z1 = std::min(nz, o[2] + extent);
        const double nd = (double)s.n;
        double av[4];
        for (int c = 0; c < 4; ++c) av[c] = (double)s.sum[c] / nd;
        double sd = 0;
        for (uint64_t z = z0; z < z1; ++z)
            for (uint64_t y = y0; y < y1; ++y) {
                const uint8_t* row = grid + 4 * (y * nx + z * nx * ny);
                for (uint64_t x = x0; x < x1; ++x) {
                    const double d0 = (double)row[4 * x + 0] - av[0], d1 = (double)row[4 * x + 1] - av[1];
                    const double d2 = (double)row[4 * x + 2] - av[2], d3 = (double)row[4 * x + 3] - av[3];
                    sd += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
                }
            }
        return std::sqrt(sd / nd);
    }

    // OctreeBuilder::insert + the stats lambda (OctreeConstruction.h:76-90, 127-139)
    uint32_t insert(const xn_node& node, bool leaf) {
        const uint32_t end_index = (uint32_t)nodes.size();
        uint32_t actual = end_index;
        if (dag) {
            auto it = cache.find(node);
            if (it != cache.end()) actual = it->second;
        }
        const bool inserted = actual == end_index;
        if (inserted) {
            if (nodes.size() >= 0xFFFFFFFFull) throw Error(XN_ERR_LIMIT, "octree exceeds 2^32 - 1 nodes");
            nodes.push_back(node);
            if (dag) cache.emplace(node, end_index);
        }
        ++stats.total_nodes;
        if (leaf) {
            ++stats.total_leaves;
            if (inserted) ++stats.unique_leaves;
        }
        return actual;
    }

    static xn_node leaf_node(uint32_t color, uint64_t depth) {
        xn_node n;
        std::memset(&n, 0, sizeof n);
        n.color = color;
        n.is_leaf_depth = LEAF | (uint32_t)depth;
        return n;
    }

    // roll the node array (and DAG cache, and stats) back to a snapshot
    struct Snapshot {
        size_t size;
        xn_build_stats stats;
    };
    Snapshot snapshot() const { return {nodes.size(), stats}; }
    void rollback(const Snapshot& s) {
        if (dag)
            for (size_t i = s.size; i < nodes.size(); ++i) cache.erase(nodes[i]);
        nodes.resize(s.size);
        stats = s.stats;
    }

    bool in_grid(const uint64_t o[3]) const { return o[0] < nx && o[1] < ny && o[2] < nz; }
    bool fully_in_grid(const uint64_t o[3], uint64_t e) const {
        return o[0] + e <= nx && o[1] + e <= ny && o[2] + e <= nz;
    }

    // detail::construct as written (OctreeConstruction.h:124-194): used for --std-dev and
    // for small blocks of the --chan-diff path
    uint32_t construct_topdown(const uint64_t o[3], uint64_t extent, uint64_t depth) {
        stats.depth = std::max<uint64_t>(stats.depth, depth);
        if (!in_grid(o)) return insert(leaf_node(0, depth), true);
        const Summary s = scan(o, extent);
        const bool split = heuristic == Heuristic::StdDev ? stddev_of(o, extent, s) > param
                                                          : s.max_diff() > (uint8_t)param;
        if ((!split && fully_in_grid(o, extent)) || extent == 1) return insert(leaf_node(s.avg_color(), depth), true);
        xn_node node;
        std::memset(&node, 0, sizeof node);
        node.color = s.avg_color();
        node.is_leaf_depth = (uint32_t)depth;
        const uint64_t h = extent / 2;
        int child = 0;
        for (int xi = 0; xi < 2; ++xi)
            for (int yi = 0; yi < 2; ++yi)
                for (int zi = 0; zi < 2; ++zi) {
                    const uint64_t co[3] = {o[0] + (xi ? h : 0), o[1] + (yi ? h : 0), o[2] + (zi ? h : 0)};
                    node.children[child++] = construct_topdown(co, h, depth + 1);
                }
        return insert(node, false);
    }

    // --chan-diff, bottom-up: emit the 8 children, then roll them back if this node is a leaf.
    // Valid because max_diff is monotone: a parent below the threshold has only children
    // below it, so the rolled-back subtrees are always 8 plain leaves.
    uint32_t construct_bottomup(const uint64_t o[3], uint64_t extent, uint64_t depth, Summary& out) {
        constexpr uint64_t BLOCK = 8;
        if (!in_grid(o)) {
            out = Summary();
            stats.depth = std::max<uint64_t>(stats.depth, depth);
            return insert(leaf_node(0, depth), true);
        }
        if (extent <= BLOCK) {
            out = scan(o, extent);
            return construct_topdown(o, extent, depth);
        }
        const Snapshot snap = snapshot();
        stats.depth = std::max<uint64_t>(stats.depth, depth);
        xn_node node;
        std::memset(&node, 0, sizeof node);
        const uint64_t h = extent / 2;
        Summary s;
        int child = 0;
        for (int xi = 0; xi < 2; ++xi)
            for (int yi = 0; yi < 2; ++yi)
                for (int zi = 0; zi < 2; ++zi) {
                    const uint64_t co[3] = {o[0] + (xi ? h : 0), o[1] + (yi ? h : 0), o[2] + (zi ? h : 0)};
                    Summary cs;
                    node.children[child++] = construct_bottomup(co, h, depth + 1, cs);
                    s.add(cs);
                }
        out = s;
        const bool split = s.max_diff() > (uint8_t)param;
        if (!split && fully_in_grid(o, extent)) {
            rollback(snap);
            stats.depth = std::max<uint64_t>(stats.depth, depth);
            return insert(leaf_node(s.avg_color(), depth), true);
        }
        node.color = s.avg_color();
        node.is_leaf_depth = (uint32_t)depth;
        return insert(node, false);
    }
};

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

// Octree::find (src/model/Octree.cpp:116-153) -> node index, 0 when pos is outside the cube
uint64_t octree_find(const std::vector<xn_node>& nodes, uint64_t dim, uint64_t px, uint64_t py, uint64_t pz,
                     uint64_t max_depth) {
    uint64_t extent = dim;
    if (px >= extent || py >= extent || pz >= extent) return 0;
    uint64_t index = 0, ox = 0, oy = 0, oz = 0;
    for (;;) {
        extent /= 2;
        if ((nodes[index].is_leaf_depth & LEAF) || extent == 0 || max_depth == 0) return index;
        uint64_t ci = 0;
        if (px >= ox + extent) { ci |= 4; ox += extent; }
        if (py >= oy + extent) { ci |= 2; oy += extent; }
        if (pz >= oz + extent) { ci |= 1; oz += extent; }
        index = nodes[index].children[ci];
        --max_depth;
    }
}

void rope_walk(std::vector<xn_node>& nodes, uint64_t dim, uint64_t px, uint64_t py, uint64_t pz, uint64_t extent,
               uint64_t depth, uint64_t index) {
    if (nodes[index].is_leaf_depth & LEAF) {
        // unsigned wrap-around of `p - extent` at the low faces is intentional: find() rejects
        // it as out of range exactly as the reference's size_t arithmetic does
        uint32_t r[6];
        r[0] = (uint32_t)octree_find(nodes, dim, px + extent, py, pz, depth);
        r[1] = (uint32_t)octree_find(nodes, dim, px - extent, py, pz, depth);
        r[2] = (uint32_t)octree_find(nodes, dim, px, py + extent, pz, depth);
        r[3] = (uint32_t)octree_find(nodes, dim, px, py - extent, pz, depth);
        r[4] = (uint32_t)octree_find(nodes, dim, px, py, pz + extent, depth);
        r[5] = (uint32_t)octree_find(nodes, dim, px, py, pz - extent, depth);
        for (int i = 0; i < 6; ++i) nodes[index].children[i] = r[i];
        return;
    }
    const uint64_t h = extent / 2;
    int child = 0;
    for (int xi = 0; xi < 2; ++xi)
        for (int yi = 0; yi < 2; ++yi)
            for (int zi = 0; zi < 2; ++zi) {
                const uint32_t ci = nodes[index].children[child++];
                rope_walk(nodes, dim, px + (xi ? h : 0), py + (yi ? h : 0), pz + (zi ? h : 0), h, depth + 1, ci);
            }
}

} // namespace

void generate_ropes(Octree& tree) {
    if (tree.nodes.empty()) return;
    rope_walk(tree.nodes, tree.side, 0, 0, 0, tree.side, 0, 0);
}

Octree build_octree(const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, Heuristic h, double param,
                    OctreeType type, xn_build_stats* stats) {
    if (!rgba || nx == 0 || ny == 0 || nz == 0) throw Error(XN_ERR_INVALID, "build_octree: empty grid");
    Builder b{rgba, nx, ny, nz, h, param, type == OctreeType::Dag, {}, {}, {0, 0, 0, 0}};
    const uint64_t dim = std::max({ceil_2pow(nx), ceil_2pow(ny), ceil_2pow(nz)});
    const uint64_t origin[3] = {0, 0, 0};
    if (h == Heuristic::ChanDiff) {
        Summary s;
        b.construct_bottomup(origin, dim, 0, s);
    } else {
        b.construct_topdown(origin, dim, 0);
    }

    // OctreeBuilder::build (OctreeConstruction.h:92-112): reverse so the root is node 0
    Octree tree;
    tree.side = dim;
    tree.nodes = std::move(b.nodes);
    tree.nodes.shrink_to_fit();
    std::reverse(tree.nodes.begin(), tree.nodes.end());
    const uint32_t end = (uint32_t)tree.nodes.size() - 1;
    for (auto& n : tree.nodes) {
        if (n.is_leaf_depth & LEAF) {
            for (auto& c : n.children) c = 0;
        } else {
            for (auto& c : n.children) c = end - c;
        }
    }
    if (type == OctreeType::Rope) generate_ropes(tree);
    if (stats) *stats = b.stats;
    return tree;
}

} // namespace xn
