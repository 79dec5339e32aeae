// xn_png.cpp -- RGBA8 PNG writer (zlib deflate), the counterpart of lodepng::encode in
// HeadlessDisplay::save (reference src/backend/headless/HeadlessDisplay.cpp:78-91).
#include <zlib.h>

#include <cstdio>
#include <cstring>

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
    out.insert(out.end(), data, data + n);
    put32(out, (uint32_t)crc32(0, out.data() + start, (uInt)(n + 4)));
}
} // namespace

void png_write(const std::string& path, const uint32_t* rgba, uint32_t w, uint32_t h) {
    if (w == 0 || h == 0) throw Error(XN_ERR_INVALID, "png_write: empty image");
    // filter type 0 (None) in front of every row
    std::vector<uint8_t> raw((size_t)h * (1 + (size_t)w * 4));
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t* row = raw.data() + (size_t)y * (1 + (size_t)w * 4);
        row[0] = 0;
        std::memcpy(row + 1, rgba + (size_t)y * w, (size_t)w * 4);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 3) != Z_OK)
        throw Error(XN_ERR_IO, "png_write: deflate failed");

    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put32(ihdr, w);
    put32(ihdr, h);
    const uint8_t tail[5] = {8, 6, 0, 0, 0}; // 8 bit, RGBA, deflate, adaptive, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
    chunk(out, "IDAT", z.data(), zlen);
    chunk(out, "IEND", nullptr, 0);

    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw Error(XN_ERR_IO, "Failed to open");
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    if (!ok) throw Error(XN_ERR_IO, "png_write: short write");
}

} // namespace xn
