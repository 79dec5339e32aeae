// xn_png.cpp -- RGBA8 PNG writer, the counterpart of lodepng::encode in HeadlessDisplay::save
// (reference src/backend/headless/HeadlessDisplay.cpp:78-91).
//
// The reference deflates every frame on one thread, which dominates wall time when frames are
// saved.  Here the rows are cut into bands that are deflated
// concurrently as raw streams ending on a byte boundary (Z_SYNC_FLUSH), and the bands are
// concatenated into one zlib stream (the last band carries the final block; the Adler-32 of the
// whole image is combined from the bands').  Any PNG reader decodes the result.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>

#include "xn_host.hpp"

namespace xn {
namespace {
void put32(std::vector<uint8_t>& v, uint32_t x) {
    for (int i = 3; i >= 0; --i) v.push_back((uint8_t)(x >> (8 * i)));
}
void chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
    put32(out, (uint32_t)n);
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put32(out, (uint32_t)crc32(0, out.data() + start, (uInt)(n + 4)));
}

struct Band {
    uint32_t row0 = 0, rows = 0;
    std::vector<uint8_t> deflated;
    uLong adler = 1;
    size_t raw_len = 0;
    bool ok = false;
};

// deflate rows [row0, row0+rows) (filter byte 0 + pixels per row) as a raw stream
void deflate_band(Band& b, const uint32_t* rgba, uint32_t w, bool last) {
    const size_t stride = 1 + (size_t)w * 4;
    std::vector<uint8_t> raw((size_t)b.rows * stride);
    for (uint32_t y = 0; y < b.rows; ++y) {
        uint8_t* row = raw.data() + (size_t)y * stride;
        row[0] = 0; // filter type None
        std::memcpy(row + 1, rgba + (size_t)(b.row0 + y) * w, (size_t)w * 4);
    }
    b.raw_len = raw.size();
    b.adler = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, 3, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return;
    b.deflated.resize(deflateBound(&zs, (uLong)raw.size()) + 16);
    zs.next_in = raw.data();
    zs.avail_in = (uInt)raw.size();
    zs.next_out = b.deflated.data();
    zs.avail_out = (uInt)b.deflated.size();
    const int rc = deflate(&zs, last ? Z_FINISH : Z_SYNC_FLUSH);
    b.ok = last ? rc == Z_STREAM_END : (rc == Z_OK && zs.avail_in == 0);
    b.deflated.resize(b.deflated.size() - zs.avail_out);
    deflateEnd(&zs);
}
} // namespace

void png_write(const std::string& path, const uint32_t* rgba, uint32_t w, uint32_t h) {
    if (w == 0 || h == 0) throw Error(XN_ERR_INVALID, "png_write: empty image");
    if ((uint64_t)(1 + (uint64_t)w * 4) * 64 > 0x7FFFFFFFull) throw Error(XN_ERR_LIMIT, "png_write: image too wide");

    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt, 32u));
    // bands of at least 64 rows, at most ~2^31 bytes each
    const uint32_t rows_per_band = std::max<uint32_t>(64, (h + nt - 1) / nt);
    std::vector<Band> bands;
    for (uint32_t y = 0; y < h; y += rows_per_band) {
        Band b;
        b.row0 = y;
        b.rows = std::min(rows_per_band, h - y);
        bands.push_back(std::move(b));
    }
    std::vector<std::thread> pool;
    for (size_t i = 0; i < bands.size(); ++i)
        pool.emplace_back([&, i] { deflate_band(bands[i], rgba, w, i + 1 == bands.size()); });
    for (auto& t : pool) t.join();

    std::vector<uint8_t> z = {0x78, 0x01}; // zlib header: deflate, 32 KiB window, no preset dictionary
    uLong adler = adler32(0L, Z_NULL, 0);
    for (const auto& b : bands) {
        if (!b.ok) throw Error(XN_ERR_IO, "png_write: deflate failed");
        z.insert(z.end(), b.deflated.begin(), b.deflated.end());
        adler = adler32_combine(adler, b.adler, (z_off_t)b.raw_len);
    }
    put32(z, (uint32_t)adler);

    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put32(ihdr, w);
    put32(ihdr, h);
    const uint8_t tail[5] = {8, 6, 0, 0, 0}; // 8 bit, RGBA, deflate, adaptive, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
    // IDAT chunks of at most 1 GiB
    for (size_t off = 0; off < z.size(); off += (1u << 30)) chunk(out, "IDAT", z.data() + off, std::min<size_t>(1u << 30, z.size() - off));
    chunk(out, "IEND", nullptr, 0);

    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw Error(XN_ERR_IO, "Failed to open");
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    if (!ok) throw Error(XN_ERR_IO, "png_write: short write");
}

} // namespace xn
