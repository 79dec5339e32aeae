// xn_tiff.cpp -- TIFF-stack volume reader/writer without libtiff.
//
// Behavioural target: Grid::load_tiff (reference src/model/Grid.cpp:27-79), i.e. what
// libtiff's TIFFReadRGBAImage(tiff, W, H, raster) leaves in memory for each directory:
//   * one directory per z slice, all slices W x H;
//   * 8-bit R,G,B,A bytes per voxel (libtiff packs ABGR into a uint32, little endian);
//   * raster row 0 is the BOTTOM row of a top-left oriented image (ORIENTATION_BOTLEFT
//     is TIFFReadRGBAImage's default), so grid y = H - 1 - file_row;
//   * unassociated alpha is pre-multiplied: c' = (c * a + 127) / 255; associated or
//     unspecified extra samples are taken as they are; without alpha, A = 255.
// Supported subset (what the reference's data tools and Pillow write): classic TIFF and
// BigTIFF, either byte order, uncompressed, chunky, strips or tiles, 8 bits per sample,
// 1 (grey), 2 (grey+alpha), 3 (RGB) or 4 (RGBA) samples per pixel.
#include <cstdio>
#include <cstring>
#include <memory>

#include "xn_host.hpp"

namespace xn {
namespace {

struct File {
    FILE* f = nullptr;
    explicit File(const std::string& path, const char* mode) : f(std::fopen(path.c_str(), mode)) {}
    ~File() {
        if (f) std::fclose(f);
    }
    File(const File&) = delete;
    File& operator=(const File&) = delete;
};

struct Reader {
    FILE* f;
    bool big_endian = false;
    bool bigtiff = false;

    void seek(uint64_t off) {
        if (fseeko(f, (off_t)off, SEEK_SET) != 0) throw Error(XN_ERR_FORMAT, "TIFF: seek past end of file");
    }
    void read(void* dst, size_t n) {
        if (std::fread(dst, 1, n, f) != n) throw Error(XN_ERR_FORMAT, "TIFF: unexpected end of file");
    }
    uint64_t uint_at(const uint8_t* p, int bytes) const {
        uint64_t v = 0;
        for (int i = 0; i < bytes; ++i) {
            const uint8_t b = big_endian ? p[i] : p[bytes - 1 - i];
            v = (v << 8) | b;
        }
        return v;
    }
    uint64_t read_uint(int bytes) {
        uint8_t b[8];
        read(b, (size_t)bytes);
        return uint_at(b, bytes);
    }
};

struct Directory {
    uint64_t width = 0, height = 0;
    uint64_t bits = 1, compression = 1, photometric = 1, samples = 1, planar = 1, orientation = 1;
    uint64_t rows_per_strip = ~0ull;
    uint64_t tile_w = 0, tile_h = 0;
    int64_t extra_sample = -1; // first ExtraSamples value, -1 = none
    std::vector<uint64_t> offsets, byte_counts; // strips or tiles
    bool tiled = false;
    uint64_t next = 0;
};

int type_size(uint64_t type) {
    switch (type) {
        case 1: case 2: case 6: case 7: return 1;
        case 3: case 8: return 2;
        case 4: case 9: case 11: case 13: return 4;
        case 5: case 10: case 12: case 16: case 17: case 18: return 8;
        default: return 0;
    }
}

std::vector<uint64_t> read_values(Reader& r, uint64_t type, uint64_t count, const uint8_t* inline_bytes, int inline_cap) {
    const int ts = type_size(type);
    if (ts == 0) return {};
    if (count > (1ull << 32)) throw Error(XN_ERR_FORMAT, "TIFF: implausible tag count");
    std::vector<uint64_t> out(count);
    const uint64_t total = (uint64_t)ts * count;
    std::vector<uint8_t> buf;
    const uint8_t* src = inline_bytes;
    if (total > (uint64_t)inline_cap) {
        const uint64_t off = r.uint_at(inline_bytes, inline_cap);
        const off_t here = ftello(r.f);
        r.seek(off);
        buf.resize(total);
        r.read(buf.data(), total);
        fseeko(r.f, here, SEEK_SET);
        src = buf.data();
    }
    // rationals etc. are never needed; read them as raw 8-byte values
    for (uint64_t i = 0; i < count; ++i) out[i] = r.uint_at(src + i * ts, ts);
    return out;
}

Directory read_directory(Reader& r, uint64_t offset) {
    Directory d;
    r.seek(offset);
    const uint64_t n = r.read_uint(r.bigtiff ? 8 : 2);
    if (n == 0 || n > 4096) throw Error(XN_ERR_FORMAT, "TIFF: implausible directory entry count");
    const int entry_size = r.bigtiff ? 20 : 12;
    std::vector<uint8_t> entries(n * entry_size);
    r.read(entries.data(), entries.size());
    d.next = r.read_uint(r.bigtiff ? 8 : 4);
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* e = entries.data() + i * entry_size;
        const uint64_t tag = r.uint_at(e, 2), type = r.uint_at(e + 2, 2);
        const uint64_t count = r.bigtiff ? r.uint_at(e + 4, 8) : r.uint_at(e + 4, 4);
        const uint8_t* val = e + (r.bigtiff ? 12 : 8);
        const int cap = r.bigtiff ? 8 : 4;
        auto scalar = [&]() -> uint64_t {
            auto v = read_values(r, type, count ? 1 : 0, val, cap);
            return v.empty() ? 0 : v[0];
        };
        switch (tag) {
            case 256: d.width = scalar(); break;
            case 257: d.height = scalar(); break;
            case 258: {
                auto v = read_values(r, type, count, val, cap);
                d.bits = v.empty() ? 1 : v[0];
                for (auto b : v)
                    if (b != d.bits) throw Error(XN_ERR_FORMAT, "TIFF: mixed bits per sample are not supported");
                break;
            }
            case 259: d.compression = scalar(); break;
            case 262: d.photometric = scalar(); break;
            case 273: d.offsets = read_values(r, type, count, val, cap); break;
            case 274: d.orientation = scalar(); break;
            case 277: d.samples = scalar(); break;
            case 278: d.rows_per_strip = scalar(); break;
            case 279: d.byte_counts = read_values(r, type, count, val, cap); break;
            case 284: d.planar = scalar(); break;
            case 322: d.tile_w = scalar(); break;
            case 323: d.tile_h = scalar(); break;
            case 324: d.offsets = read_values(r, type, count, val, cap); d.tiled = true; break;
            case 325: d.byte_counts = read_values(r, type, count, val, cap); break;
            case 338: {
                auto v = read_values(r, type, count, val, cap);
                if (!v.empty()) d.extra_sample = (int64_t)v[0];
                break;
            }
            default: break;
        }
    }
    return d;
}

Reader open_reader(FILE* f, uint64_t& first_dir) {
    Reader r{f};
    uint8_t hdr[16];
    r.read(hdr, 4);
    if (hdr[0] == 'I' && hdr[1] == 'I') r.big_endian = false;
    else if (hdr[0] == 'M' && hdr[1] == 'M') r.big_endian = true;
    else throw Error(XN_ERR_FORMAT, "Failed to open"); // what a non-TIFF yields in Grid::load_tiff
    const uint64_t magic = r.uint_at(hdr + 2, 2);
    if (magic == 42) {
        first_dir = r.read_uint(4);
    } else if (magic == 43) {
        r.bigtiff = true;
        r.read(hdr, 4); // offset size (8), reserved (0)
        if (r.uint_at(hdr, 2) != 8) throw Error(XN_ERR_FORMAT, "TIFF: unsupported BigTIFF offset size");
        first_dir = r.read_uint(8);
    } else {
        throw Error(XN_ERR_FORMAT, "Failed to open");
    }
    return r;
}

std::vector<Directory> read_directories(Reader& r, uint64_t first) {
    std::vector<Directory> dirs;
    uint64_t off = first;
    while (off != 0) {
        dirs.push_back(read_directory(r, off));
        off = dirs.back().next;
        if (dirs.size() > (1u << 20)) throw Error(XN_ERR_FORMAT, "TIFF: directory chain too long");
    }
    if (dirs.empty()) throw Error(XN_ERR_FORMAT, "TIFF: no image directories");
    return dirs;
}

// Grid::load_tiff's dimension pass (src/model/Grid.cpp:37-57)
TiffInfo check_dims(const std::vector<Directory>& dirs) {
    uint64_t width = 0, height = 0;
    uint64_t depth = 0;
    for (const auto& d : dirs) {
        if (d.width == 0 || d.height == 0)
            throw Error(XN_ERR_FORMAT, "Layer " + std::to_string(depth) + " has invalid dimensions (" +
                                           std::to_string(d.width) + "x" + std::to_string(d.height) + ")");
        if (width == 0) {
            width = d.width;
            height = d.height;
        } else if (d.width != width || d.height != height) {
            // The reference only rejects a layer whose width AND height both differ (`&&`,
            // Grid.cpp:47) and then reads a mismatched raster; every mismatch is rejected here.
            throw Error(XN_ERR_FORMAT, "Dimensions of layer " + std::to_string(depth) +
                                           " differ from previous dimensions (" + std::to_string(d.width) + "x" +
                                           std::to_string(d.height) + ", previously " + std::to_string(width) + "x" +
                                           std::to_string(height) + ")");
        }
        ++depth;
    }
    // the reference counts layers in a uint16_t (Grid.cpp:35)
    if (depth > 65535) throw Error(XN_ERR_LIMIT, "TIFF: more than 65535 layers");
    // libtiff's own limit on a raster is uint32 per axis; here the grid's byte size must also be
    // representable, so that no later size computation (layer = W*H*4, layer*depth, strip and
    // tile byte counts, all bounded by it) can wrap on a crafted file
    uint64_t bytes = 0;
    if (width > 0x7FFFFFFFull || height > 0x7FFFFFFFull || __builtin_mul_overflow(width, height, &bytes) ||
        __builtin_mul_overflow(bytes, (uint64_t)4, &bytes) || __builtin_mul_overflow(bytes, depth, &bytes) ||
        bytes > (1ull << 62))
        throw Error(XN_ERR_LIMIT, "TIFF: volume of " + std::to_string(width) + "x" + std::to_string(height) + "x" +
                                      std::to_string(depth) + " voxels is too large");
    return TiffInfo{width, height, depth};
}

// Which sample is alpha and whether colours get pre-multiplied, as libtiff's TIFFRGBAImage decides
// (tif_getimage.c, TIFFRGBAImageBegin + PickContigCase; pinned against libtiff 4.7 in the tests):
//   RGB : a 4th sample is alpha; unassociated (ExtraSamples = 2) is pre-multiplied, associated or
//         unspecified is taken as it is;
//   grey: only a 2-sample image whose ExtraSamples says associated (1) or unassociated (2) has alpha,
//         and its grey value is never pre-multiplied (putagreytile); anything else is opaque.
void alpha_rules(const Directory& d, bool& has_alpha, bool& unassociated) {
    if (d.photometric == 2) {
        has_alpha = d.samples > 3;
        unassociated = has_alpha && d.extra_sample == 2;
    } else {
        has_alpha = d.samples == 2 && (d.extra_sample == 1 || d.extra_sample == 2);
        unassociated = false;
    }
}

// What TIFFReadRGBAImage does with the file's Orientation tag when asked for its default
// bottom-left raster (libtiff tif_getimage.c, setorientation): bit 0 = rows are reversed, bit 1 =
// columns are reversed.  Top-left / left-top (1, 5): rows; top-right / right-top (2, 6): both;
// bottom-right / right-bottom (3, 7): columns; bottom-left / left-bottom (4, 8): neither.  (libtiff
// does not transpose the 5-8 orientations either.)  Unknown values are read as the default, 1.
uint32_t orientation_flips(uint64_t orientation) {
    switch (orientation) {
        case 2: case 6: return 3u;
        case 3: case 7: return 2u;
        case 4: case 8: return 0u;
        default: return 1u;
    }
}

void decode_directory(Reader& r, const Directory& d, uint8_t* out /* W*H*4, grid row order */) {
    if (d.compression != 1) throw Error(XN_ERR_FORMAT, "TIFF: only uncompressed data is supported");
    if (d.bits != 8) throw Error(XN_ERR_FORMAT, "TIFF: only 8 bits per sample are supported");
    if (d.planar != 1 && d.samples > 1) throw Error(XN_ERR_FORMAT, "TIFF: planar configuration 2 is not supported");
    if (d.samples < 1 || d.samples > 4) throw Error(XN_ERR_FORMAT, "TIFF: unsupported samples per pixel");
    if (d.photometric > 2) throw Error(XN_ERR_FORMAT, "TIFF: unsupported photometric interpretation");
    const uint64_t W = d.width, H = d.height, spp = d.samples;
    const bool rgb = d.photometric == 2;
    if (rgb && spp < 3) throw Error(XN_ERR_FORMAT, "TIFF: RGB image with fewer than 3 samples");
    const uint64_t colour_samples = rgb ? 3 : 1;
    bool has_alpha, unassociated;
    alpha_rules(d, has_alpha, unassociated);
    const uint32_t flips = orientation_flips(d.orientation);
    const bool flip = (flips & 1u) != 0u, mirror = (flips & 2u) != 0u;

    auto put = [&](uint64_t x, uint64_t file_row, const uint8_t* px) {
        uint8_t rr, gg, bb, aa = 255;
        if (rgb) {
            rr = px[0]; gg = px[1]; bb = px[2];
        } else {
            const uint8_t v = d.photometric == 0 ? (uint8_t)(255 - px[0]) : px[0];
            rr = gg = bb = v;
        }
        if (has_alpha) {
            aa = px[colour_samples];
            if (unassociated) {
                rr = (uint8_t)((rr * aa + 127) / 255);
                gg = (uint8_t)((gg * aa + 127) / 255);
                bb = (uint8_t)((bb * aa + 127) / 255);
            }
        }
        const uint64_t y = flip ? H - 1 - file_row : file_row;
        const uint64_t xo = mirror ? W - 1 - x : x;
        uint8_t* o = out + 4 * (xo + y * W);
        o[0] = rr; o[1] = gg; o[2] = bb; o[3] = aa;
    };

    std::vector<uint8_t> buf;
    if (!d.tiled) {
        const uint64_t rps = d.rows_per_strip < H ? d.rows_per_strip : H;
        if (rps == 0) throw Error(XN_ERR_FORMAT, "TIFF: zero rows per strip");
        const uint64_t nstrips = (H + rps - 1) / rps;
        if (d.offsets.size() < nstrips) throw Error(XN_ERR_FORMAT, "TIFF: missing strip offsets");
        for (uint64_t s = 0; s < nstrips; ++s) {
            const uint64_t row0 = s * rps, rows = (row0 + rps <= H) ? rps : H - row0;
            const uint64_t bytes = rows * W * spp;
            if (!d.byte_counts.empty() && s < d.byte_counts.size() && d.byte_counts[s] < bytes)
                throw Error(XN_ERR_FORMAT, "TIFF: strip is shorter than its rows");
            buf.resize(bytes);
            r.seek(d.offsets[s]);
            r.read(buf.data(), bytes);
            for (uint64_t rr = 0; rr < rows; ++rr)
                for (uint64_t x = 0; x < W; ++x) put(x, row0 + rr, buf.data() + (rr * W + x) * spp);
        }
    } else {
        if (d.tile_w == 0 || d.tile_h == 0) throw Error(XN_ERR_FORMAT, "TIFF: zero tile size");
        const uint64_t tx = (W + d.tile_w - 1) / d.tile_w, ty = (H + d.tile_h - 1) / d.tile_h;
        if (d.offsets.size() < tx * ty) throw Error(XN_ERR_FORMAT, "TIFF: missing tile offsets");
        uint64_t bytes = 0;
        if (d.tile_w > 0x7FFFFFFFull || d.tile_h > 0x7FFFFFFFull || __builtin_mul_overflow(d.tile_w, d.tile_h, &bytes) ||
            __builtin_mul_overflow(bytes, spp, &bytes) || bytes > (1ull << 40))
            throw Error(XN_ERR_LIMIT, "TIFF: tile size too large");
        buf.resize(bytes);
        for (uint64_t j = 0; j < ty; ++j)
            for (uint64_t i = 0; i < tx; ++i) {
                r.seek(d.offsets[j * tx + i]);
                r.read(buf.data(), bytes);
                for (uint64_t yy = 0; yy < d.tile_h && j * d.tile_h + yy < H; ++yy)
                    for (uint64_t xx = 0; xx < d.tile_w && i * d.tile_w + xx < W; ++xx)
                        put(i * d.tile_w + xx, j * d.tile_h + yy, buf.data() + (yy * d.tile_w + xx) * spp);
            }
    }
}

} // namespace

TiffInfo tiff_info(const std::string& path) {
    File file(path, "rb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    uint64_t first;
    Reader r = open_reader(file.f, first);
    return check_dims(read_directories(r, first));
}

void tiff_read(const std::string& path, uint8_t* out, uint64_t cap_bytes) {
    File file(path, "rb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    uint64_t first;
    Reader r = open_reader(file.f, first);
    const auto dirs = read_directories(r, first);
    const TiffInfo info = check_dims(dirs);
    const uint64_t layer = info.nx * info.ny * 4;
    if (cap_bytes < layer * info.nz) throw Error(XN_ERR_INVALID, "tiff_read: output buffer too small");
    for (uint64_t z = 0; z < info.nz; ++z) decode_directory(r, dirs[z], out + z * layer);
}

TiffPlan tiff_plan(const std::string& path) {
    File file(path, "rb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    uint64_t first;
    Reader r = open_reader(file.f, first);
    const auto dirs = read_directories(r, first);
    TiffPlan plan;
    plan.info = check_dims(dirs);
    plan.streamable = true;
    plan.slices.resize(plan.info.nz);
    for (uint64_t z = 0; z < plan.info.nz && plan.streamable; ++z) {
        const Directory& d = dirs[z];
        const bool rgb = d.photometric == 2;
        if (d.compression != 1 || d.bits != 8 || d.tiled || d.samples < 1 || d.samples > 4 || d.photometric > 2 ||
            (d.planar != 1 && d.samples > 1) || (rgb && d.samples < 3)) {
            plan.streamable = false; // the host decoder reports what is unsupported, or handles tiles
            break;
        }
        TiffSliceFormat f{};
        f.samples = (uint32_t)d.samples;
        f.photometric = (uint32_t)d.photometric;
        bool has_alpha, unassociated;
        alpha_rules(d, has_alpha, unassociated);
        f.has_alpha = has_alpha;
        f.unassociated = unassociated;
        f.flip = orientation_flips(d.orientation);
        if (z == 0) plan.format = f;
        else if (std::memcmp(&f, &plan.format, sizeof f) != 0) plan.streamable = false;
        const uint64_t W = d.width, H = d.height;
        const uint64_t rps = d.rows_per_strip < H ? d.rows_per_strip : H;
        if (rps == 0) { plan.streamable = false; break; }
        const uint64_t nstrips = (H + rps - 1) / rps;
        if (d.offsets.size() < nstrips) { plan.streamable = false; break; }
        auto& runs = plan.slices[z];
        for (uint64_t s = 0; s < nstrips; ++s) {
            const uint64_t row0 = s * rps, rows = (row0 + rps <= H) ? rps : H - row0;
            const uint64_t bytes = rows * W * d.samples;
            if (!d.byte_counts.empty() && s < d.byte_counts.size() && d.byte_counts[s] < bytes) {
                plan.streamable = false;
                break;
            }
            if (!runs.empty() && runs.back().offset + runs.back().bytes == d.offsets[s]) runs.back().bytes += bytes;
            else runs.push_back(TiffRun{d.offsets[s], bytes});
        }
    }
    if (!plan.streamable) plan.slices.clear();
    return plan;
}

void tiff_read_slice(const std::string& path, uint64_t z, uint8_t* out) {
    File file(path, "rb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    uint64_t first;
    Reader r = open_reader(file.f, first);
    const auto dirs = read_directories(r, first);
    check_dims(dirs);
    if (z >= dirs.size()) throw Error(XN_ERR_INVALID, "tiff_read_slice: no such layer");
    decode_directory(r, dirs[z], out);
}

Grid load_tiff(const std::string& path) {
    const TiffInfo info = tiff_info(path);
    Grid g;
    g.nx = info.nx;
    g.ny = info.ny;
    g.nz = info.nz;
    g.rgba.resize(4 * g.voxels());
    tiff_read(path, g.rgba.data(), g.rgba.size());
    return g;
}

// ---- writer: one uncompressed RGBA strip per directory, rows stored top-down so that
// reading it back (bottom-up raster) reproduces `rgba` exactly ----
void tiff_write(const std::string& path, const uint8_t* rgba, uint64_t nx, uint64_t ny, uint64_t nz, bool bigtiff) {
    if (nx == 0 || ny == 0 || nz == 0) throw Error(XN_ERR_INVALID, "tiff_write: empty volume");
    const uint64_t layer = nx * ny * 4;
    if (!bigtiff && layer * nz + nz * 256 + 16 >= (1ull << 32))
        throw Error(XN_ERR_LIMIT, "tiff_write: volume too large for classic TIFF, use BigTIFF");
    File file(path, "wb");
    if (!file.f) throw Error(XN_ERR_IO, "Failed to open");
    FILE* f = file.f;
    auto w16 = [&](uint64_t v) { uint8_t b[2] = {(uint8_t)v, (uint8_t)(v >> 8)}; std::fwrite(b, 1, 2, f); };
    auto w32 = [&](uint64_t v) { uint8_t b[4]; for (int i = 0; i < 4; ++i) b[i] = (uint8_t)(v >> (8 * i)); std::fwrite(b, 1, 4, f); };
    auto w64 = [&](uint64_t v) { uint8_t b[8]; for (int i = 0; i < 8; ++i) b[i] = (uint8_t)(v >> (8 * i)); std::fwrite(b, 1, 8, f); };
    auto woff = [&](uint64_t v) { bigtiff ? w64(v) : w32(v); };

    // header
    std::fwrite("II", 1, 2, f);
    const uint64_t header = bigtiff ? 16 : 8;
    if (bigtiff) { w16(43); w16(8); w16(0); w64(header); } else { w16(42); w32(header); }

    const int n_entries = 11;
    const uint64_t ifd_size = bigtiff ? 8 + 20 * n_entries + 8 : 2 + 12 * n_entries + 4;
    const uint64_t bits_size = 8; // 4 shorts, out of line
    const uint64_t block = ifd_size + bits_size + layer;
    std::vector<uint8_t> rowbuf(nx * 4);
    for (uint64_t z = 0; z < nz; ++z) {
        const uint64_t base = header + z * block;
        const uint64_t bits_off = base + ifd_size, data_off = bits_off + bits_size;
        auto entry = [&](uint64_t tag, uint64_t type, uint64_t count, uint64_t value) {
            w16(tag); w16(type);
            if (bigtiff) { w64(count); w64(value); } else { w32(count); w32(value); }
        };
        if (bigtiff) w64(n_entries); else w16(n_entries);
        entry(256, 4, 1, nx);         // ImageWidth
        entry(257, 4, 1, ny);         // ImageLength
        entry(258, 3, 4, bigtiff ? (8ull | 8ull << 16 | 8ull << 32 | 8ull << 48) : bits_off); // BitsPerSample
        entry(259, 3, 1, 1);          // Compression: none
        entry(262, 3, 1, 2);          // Photometric: RGB
        entry(273, bigtiff ? 16 : 4, 1, data_off); // StripOffsets
        entry(277, 3, 1, 4);          // SamplesPerPixel
        entry(278, 4, 1, ny);         // RowsPerStrip
        entry(279, bigtiff ? 16 : 4, 1, layer);    // StripByteCounts
        entry(284, 3, 1, 1);          // PlanarConfiguration: chunky
        entry(338, 3, 1, 1);          // ExtraSamples: associated alpha, so a read returns the bytes unchanged
        woff(z + 1 < nz ? base + block : 0);
        w16(8); w16(8); w16(8); w16(8);
        for (uint64_t row = 0; row < ny; ++row) {
            const uint8_t* src = rgba + 4 * ((ny - 1 - row) * nx + z * nx * ny);
            if (std::fwrite(src, 1, nx * 4, f) != nx * 4) throw Error(XN_ERR_IO, "tiff_write: short write");
        }
    }
    if (std::fflush(f) != 0) throw Error(XN_ERR_IO, "tiff_write: flush failed");
}

} // namespace xn
