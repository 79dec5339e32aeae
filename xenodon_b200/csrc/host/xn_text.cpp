// xn_text.cpp -- the small text formats of the path: headless.conf, camera scripts, stats files.
#include <cctype>
#include <charconv>
#include <cstdio>
#include <limits>
#include <optional>
#include <sstream>

#include "xn_host.hpp"

namespace xn {

// ---------------------------------------------------------------------------------
// headless.conf: `device { vkindex = N  offset = (x, y)  extent = (w, h) }` blocks.
// Grammar and error texts follow the reference's recursive-descent parser
// (src/core/Parser.cpp, src/core/Config.h:169-240, HeadlessConfig.cpp:5-28): alphabetic keys,
// whitespace-separated items, keys of a block in any order, '#' rejected, unsigned decimals.
// ---------------------------------------------------------------------------------
namespace {

struct ConfParser {
    const std::string& s;
    size_t i = 0;
    size_t line = 1, column = 1;

    [[noreturn]] void fail(const std::string& msg) const {
        throw Error(XN_ERR_FORMAT,
                    "Parse error at line " + std::to_string(line) + ", col " + std::to_string(column) + ": " + msg);
    }
    int peek() const {
        if (i >= s.size()) return -1;
        const int c = (unsigned char)s[i];
        if (c == '#') fail("Unexpected character '#'");
        return c;
    }
    int consume() {
        if (i >= s.size()) fail("Unexpected end of input");
        const int c = (unsigned char)s[i++];
        if (c == '\n') {
            ++line;
            column = 0;
        }
        ++column;
        return c;
    }
    void expect(int expected) {
        const int actual = consume();
        if (actual != expected)
            fail(std::string("Expected character '") + (char)expected + "', found '" + (char)actual + "'");
    }
    void optws() {
        while (peek() >= 0 && std::isspace(peek())) consume();
    }
    std::string key() {
        const int c = consume();
        if (!std::isalpha(c)) fail(std::string("Expected alphabetic key, found '") + (char)c + "'");
        std::string k(1, (char)c);
        while (peek() >= 0 && std::isalpha(peek())) k.push_back((char)consume());
        return k;
    }
    uint64_t number() {
        int c = consume();
        if (c < '0' || c > '9') fail(std::string("Expected numeric character, found '") + (char)c + "'");
        uint64_t v = (uint64_t)(c - '0');
        c = peek();
        while (c >= '0' && c <= '9') {
            v = v * 10 + (uint64_t)(c - '0');
            consume();
            c = peek();
        }
        return v;
    }
    std::pair<uint64_t, uint64_t> pair() {
        expect('(');
        optws();
        const uint64_t x = number();
        optws();
        expect(',');
        optws();
        const uint64_t y = number();
        optws();
        expect(')');
        return {x, y};
    }
};

[[noreturn]] void config_error(const std::string& msg) { throw Error(XN_ERR_FORMAT, "Configuration error: " + msg); }

xn_headless_device parse_device_block(ConfParser& p) {
    std::optional<uint64_t> index;
    std::optional<std::pair<uint64_t, uint64_t>> offset, extent;
    p.optws();
    int c = p.peek();
    while (c >= 0 && std::isalpha(c)) {
        const std::string k = p.key();
        p.optws();
        if (k == "vkindex") {
            p.expect('=');
            p.optws();
            const uint64_t v = p.number();
            if (index) config_error("Ambiguous key 'vkindex'");
            index = v;
        } else if (k == "offset") {
            p.expect('=');
            p.optws();
            const auto v = p.pair();
            if (offset) config_error("Ambiguous key 'offset'");
            offset = v;
        } else if (k == "extent") {
            p.expect('=');
            p.optws();
            const auto v = p.pair();
            if (extent) config_error("Ambiguous key 'extent'");
            extent = v;
        } else {
            p.fail("Unexpected key '" + k + "'");
        }
        c = p.peek();
        if (c < 0 || !std::isspace(c)) break;
        p.optws();
        c = p.peek();
    }
    p.optws();
    if (!index) config_error("Missing key 'vkindex'");
    if (!offset) config_error("Missing key 'offset'");
    if (!extent) config_error("Missing key 'extent'");
    xn_headless_device d;
    d.vkindex = (uint32_t)*index;
    d.region.x = (int32_t)offset->first;
    d.region.y = (int32_t)offset->second;
    d.region.w = (uint32_t)extent->first;
    d.region.h = (uint32_t)extent->second;
    return d;
}

} // namespace

std::vector<xn_headless_device> parse_headless_config(const std::string& text) {
    ConfParser p{text};
    std::vector<xn_headless_device> devices;
    p.optws();
    int c = p.peek();
    while (c >= 0 && std::isalpha(c)) {
        const std::string k = p.key();
        p.optws();
        if (k != "device") p.fail("Unexpected key '" + k + "'");
        p.expect('{');
        devices.push_back(parse_device_block(p));
        p.expect('}');
        c = p.peek();
        if (c < 0 || !std::isspace(c)) break;
        p.optws();
        c = p.peek();
    }
    p.optws();
    c = p.peek();
    if (c > 0) p.fail(std::string("Expected end of input, found '") + (char)c + "'");
    if (devices.empty()) config_error("At least one device entry is required");
    return devices;
}

xn_rect rect_union(const xn_rect& a, const xn_rect& b) {
    const int32_t x = std::min(a.x, b.x), y = std::min(a.y, b.y);
    const uint32_t w = std::max((uint32_t)a.x + a.w, (uint32_t)b.x + b.w) - (uint32_t)x;
    const uint32_t h = std::max((uint32_t)a.y + a.h, (uint32_t)b.y + b.h) - (uint32_t)y;
    return xn_rect{x, y, w, h};
}

// ---------------------------------------------------------------------------------
// camera script: 9 whitespace-separated floats per frame, parsed like `istream >> float`;
// the stream ends the run at EOF after `>> std::ws` (ScriptCameraController.cpp:17-41)
// ---------------------------------------------------------------------------------
std::vector<CameraFrame> parse_camera_script(const std::string& text) {
    std::istringstream in(text);
    std::vector<CameraFrame> frames;
    for (;;) {
        CameraFrame f;
        float* dst[9] = {&f.forward[0], &f.forward[1], &f.forward[2], &f.up[0],         &f.up[1],
                         &f.up[2],      &f.translation[0], &f.translation[1], &f.translation[2]};
        for (float* d : dst) in >> *d;
        if (in.fail()) throw Error(XN_ERR_FORMAT, "Syntax error in camera input file");
        frames.push_back(f);
        in >> std::ws;
        if (in.eof()) break;
    }
    return frames;
}

// ---------------------------------------------------------------------------------
// stats
// ---------------------------------------------------------------------------------
void stats_combine(xn_render_stats& into, const xn_render_stats& o) {
    into.total_rays += o.total_rays;
    into.outputs += o.outputs;
    into.total_render_time += o.total_render_time;
    into.max_render_time = std::max(into.max_render_time, o.max_render_time);
    into.min_render_time = std::min(into.min_render_time, o.min_render_time);
}

double stats_mrays_per_s(const xn_render_stats& s) { return (double)s.total_rays / (s.total_render_time * 1000.0); }

// fmt's `{}` for a double is the shortest representation that round-trips
std::string fmt_double(double v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    return std::string(buf, r.ptr);
}

std::string stats_format(const std::vector<xn_render_stats>& frames, double wall_seconds) {
    uint64_t rays = 0;
    double ms = 0;
    for (const auto& f : frames) {
        rays += f.total_rays;
        ms += f.total_render_time;
    }
    std::string out;
    out += "total rays: " + std::to_string(rays) + "\n";
    out += "total render time: " + fmt_double(ms) + "\n";
    out += "total mray/s: " + fmt_double((double)rays / (ms * 1000.0)) + "\n";
    out += "average fps: " + fmt_double((double)frames.size() / wall_seconds) + "\n";
    out += "frames: " + std::to_string(frames.size()) + "\n";
    out += "# Frame number: total rays, outputs, total render time, max render time, min render time, mray/s\n";
    for (size_t i = 0; i < frames.size(); ++i) {
        const auto& f = frames[i];
        out += "frame " + std::to_string(i) + ": " + std::to_string(f.total_rays) + " rays, " +
               std::to_string(f.outputs) + ", " + fmt_double(f.total_render_time) + " ms, " +
               fmt_double(f.max_render_time) + " ms, " + fmt_double(f.min_render_time) + " ms, " +
               fmt_double(stats_mrays_per_s(f)) + " mray/s\n";
    }
    return out;
}

void stats_write(const std::string& path, const xn_render_stats* frames, uint64_t n, double wall_seconds) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw Error(XN_ERR_IO, "Failed to open render stats output path '" + path + "'");
    const std::string text = stats_format(std::vector<xn_render_stats>(frames, frames + n), wall_seconds);
    const bool ok = std::fwrite(text.data(), 1, text.size(), f) == text.size();
    std::fclose(f);
    if (!ok) throw Error(XN_ERR_IO, "Failed to write render stats");
}

} // namespace xn
