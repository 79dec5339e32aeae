// xn_cli.cpp -- the `xenodon` command line on top of the C ABI (include/xenodon_b200.h).
//
// Keeps the reference's drop-in surface for the traversal path (reference src/main.cpp,
// src/main_loop.cpp, src/convert.cpp): subcommands help / sysinfo / render / convert, the
// render and convert flag sets, the log lines, the frame-loop semantics (--repeat, camera
// script EOF ends the run, --discard-output means no readback) and the stats file.
// Presentation back ends (--xorg, --direct) are not part of this path and report an error.
#include <charconv>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <limits>
#include <sstream>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "xenodon_b200.h"

namespace {

struct CliError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] void lib_fail() { throw CliError(xn_last_error()); }
void check(int rc) {
    if (rc != XN_OK) lib_fail();
}

std::string fmt_double(double v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    return std::string(buf, r.ptr);
}

// ---- logger: "[HH:MM:SS] message" to the console and/or a file (reference src/core/Logger.cpp) ----
struct Logger {
    bool console = false;
    std::ofstream file;
    void log(const std::string& msg) {
        char stamp[16];
        std::time_t t = std::time(nullptr);
        std::tm tmv;
        localtime_r(&t, &tmv);
        std::strftime(stamp, sizeof stamp, "%H:%M:%S", &tmv);
        const std::string line = std::string("[") + stamp + "] " + msg + "\n";
        static std::mutex mu; // the frame writer logs from its own thread
        std::lock_guard<std::mutex> lock(mu);
        if (console) {
            std::fputs(line.c_str(), stdout);
            std::fflush(stdout);
        }
        if (file.is_open()) {
            file << line;
            file.flush();
        }
    }
} LOGGER;

// ---- command-line table ----
// One table per subcommand: switches (no value), options (one value) and operands, looked up by
// spelling in a map built once.  The accepted spellings and the error texts are the reference's
// command-line surface (src/core/arg_parse.cpp); the mechanism is this file's own.
using Setter = std::function<bool(const char*)>;
class ArgTable {
    struct Entry {
        std::string spelling; // "--quiet"
        char letter;          // 'q' or 0
        std::string value;    // name of the value for messages; empty for a switch
        Setter set;           // switches: called with nullptr
        bool used = false;
        std::string shown() const { return letter ? spelling + "/-" + letter : spelling; }
    };
    std::vector<Entry> entries;
    std::vector<std::pair<std::string, Setter>> operands;

public:
    ArgTable& toggle(const char* spelling, char letter, bool* target) {
        entries.push_back({spelling, letter, "", [target](const char*) { *target = true; return true; }});
        return *this;
    }
    ArgTable& option(const char* spelling, char letter, const char* value, Setter set) {
        entries.push_back({spelling, letter, value, std::move(set)});
        return *this;
    }
    ArgTable& operand(const char* name, Setter set) {
        operands.emplace_back(name, std::move(set));
        return *this;
    }
    void apply(const std::vector<const char*>& words) {
        std::unordered_map<std::string, Entry*> by_spelling;
        for (auto& e : entries) {
            by_spelling[e.spelling] = &e;
            if (e.letter) by_spelling[std::string("-") + e.letter] = &e;
        }
        size_t next_operand = 0;
        for (auto w = words.begin(); w != words.end(); ++w) {
            const std::string word = *w;
            if (word.empty() || word[0] != '-') {
                if (next_operand == operands.size()) throw CliError("Unexpected positional argument '" + word + "'");
                auto& [name, set] = operands[next_operand++];
                if (!set(*w)) throw CliError("Invalid value for positional argument <" + name + ">");
                continue;
            }
            const auto hit = by_spelling.find(word);
            if (hit == by_spelling.end()) throw CliError("Unrecognized option " + word);
            Entry& e = *hit->second;
            const bool takes_value = !e.value.empty();
            if (e.used)
                throw CliError(std::string("Duplicate specification of ") + (takes_value ? "parameter " : "flag ") + e.shown());
            e.used = true;
            if (!takes_value) {
                e.set(nullptr);
            } else if (++w == words.end()) {
                throw CliError("Parameter " + word + " expects argument <" + e.value + ">");
            } else if (!e.set(*w)) {
                throw CliError("Invalid value for <" + e.value + "> of parameter " + word);
            }
        }
        if (next_operand != operands.size())
            throw CliError("Missing required positional argument <" + operands[next_operand].first + ">");
    }
};
using Action = Setter;

// digits, '.' and '-' only (no exponent), src/core/arg_parse.h:75-100
template <typename T>
bool parse_float(std::string_view s, T& out) {
    for (char c : s)
        if (!std::isdigit((unsigned char)c) && c != '.' && c != '-') return false;
    const std::string tmp(s);
    if (tmp.empty()) return false;
    char* end = nullptr;
    const double v = std::strtod(tmp.c_str(), &end);
    if (end != tmp.c_str() + tmp.size()) return false;
    out = std::is_same_v<T, float> ? (T)std::strtof(tmp.c_str(), nullptr) : (T)v;
    return true;
}
Action string_opt(std::string* var) {
    return [var](const char* a) { *var = a; return true; };
}
template <typename T>
Action float_min_opt(T* var, T min) {
    return [var, min](const char* a) {
        T v;
        if (!parse_float<T>(a, v) || v < min) return false;
        *var = v;
        return true;
    };
}
template <typename T>
Action int_range_opt(T* var, T min, T max) {
    return [var, min, max](const char* a) {
        const std::string_view s = a;
        T v;
        auto [end, err] = std::from_chars(s.data(), s.data() + s.size(), v);
        if (err != std::errc() || end != s.data() + s.size() || v < min || v > max) return false;
        *var = v;
        return true;
    };
}

// ---- fmt-style output name: `out-{}.png`, `out-{:0>3}.png`, `{0}`, `{:03}`, `{{` ----
std::string format_frame_name(const std::string& pattern, uint64_t frame) {
    std::string out;
    for (size_t i = 0; i < pattern.size(); ++i) {
        const char c = pattern[i];
        if (c == '{') {
            if (i + 1 < pattern.size() && pattern[i + 1] == '{') { out += '{'; ++i; continue; }
            const size_t close = pattern.find('}', i);
            if (close == std::string::npos) throw CliError("Failed to format output filename: invalid format string");
            std::string spec = pattern.substr(i + 1, close - i - 1);
            i = close;
            std::string fmt;
            const size_t colon = spec.find(':');
            const std::string index = colon == std::string::npos ? spec : spec.substr(0, colon);
            if (colon != std::string::npos) fmt = spec.substr(colon + 1);
            if (!index.empty() && index != "0")
                throw CliError("Failed to format output filename: argument index out of range");
            char fill = ' ', align = '>';
            size_t k = 0;
            if (fmt.size() >= 2 && (fmt[1] == '<' || fmt[1] == '>' || fmt[1] == '^')) { fill = fmt[0]; align = fmt[1]; k = 2; }
            else if (!fmt.empty() && (fmt[0] == '<' || fmt[0] == '>' || fmt[0] == '^')) { align = fmt[0]; k = 1; }
            if (k < fmt.size() && fmt[k] == '0') { fill = '0'; align = '>'; ++k; }
            size_t width = 0;
            while (k < fmt.size() && std::isdigit((unsigned char)fmt[k])) width = width * 10 + (size_t)(fmt[k++] - '0');
            if (k < fmt.size() && fmt[k] == 'd') ++k;
            if (k != fmt.size()) throw CliError("Failed to format output filename: invalid format specifier");
            std::string num = std::to_string(frame);
            if (num.size() < width) {
                const size_t pad = width - num.size();
                if (align == '<') num += std::string(pad, fill);
                else if (align == '^') num = std::string(pad / 2, fill) + num + std::string(pad - pad / 2, fill);
                else num = std::string(pad, fill) + num;
            }
            out += num;
        } else if (c == '}') {
            if (i + 1 < pattern.size() && pattern[i + 1] == '}') { out += '}'; ++i; continue; }
            throw CliError("Failed to format output filename: unmatched '}' in format string");
        } else {
            out += c;
        }
    }
    return out;
}

std::string read_text_file(const std::string& path, const char* what) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw CliError(std::string(what) + " '" + path + "'");
    std::ostringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

// ---------------------------------------------------------------------------------
// help
// ---------------------------------------------------------------------------------
const char* HELP_MAIN =
    "Usage: xenodon <subcommand> [options...]\n"
    "\n"
    "Volumetric ray tracer: B200-native (CUDA, sm_100a) implementation of Xenodon's traversal path.\n"
    "\n"
    "Subcommands:\n"
    "  help [topic]   Show this text, or help on: help, sysinfo, convert, render, headless-config\n"
    "  sysinfo        List the CUDA devices usable as 'vkindex' in a headless configuration\n"
    "  render         Render a volume (tiff grid or svo octree) with the headless backend\n"
    "  convert        Convert a tiff volume into a sparse voxel octree (.svo)\n";
const char* HELP_RENDER =
    "Usage: xenodon render [options...] <volume path>\n"
    "\n"
    "Backend (exactly one is required; only the headless backend is part of this build):\n"
    "  --headless <config path>     Render off-screen to the regions described by the configuration\n"
    "                               (see 'xenodon help headless-config').\n"
    "  --output <output path>       fmt-style pattern for saved frames, the frame number is argument 0.\n"
    "                               Default 'out-{}.png'.\n"
    "  --discard-output             Do not read frames back or save them (benchmarking).\n"
    "  --xorg, --xorg-multi-gpu <config path>, --direct <config path>\n"
    "                               Presentation backends of the reference; not available here.\n"
    "\n"
    "Options:\n"
    "  -q, --quiet                  Do not log to the console.\n"
    "  --log-output <output path>   Also write the log to this file.\n"
    "  -e, --emission-coeff <f>     Emission coefficient (>= 0, default 1).\n"
    "  --volume-type <tiff|tif|svo> Override the volume type guessed from the file extension.\n"
    "  -s, --shader <name>          Traversal: dda (tiff); svo-naive, esvo, svo-df, svo-rope (svo).\n"
    "                               Default: dda for tiff volumes, svo-naive for svo volumes.\n"
    "  -r, --voxel-ratio <x:y:z>    Physical size ratio of a voxel (each > 0, default 1:1:1).\n"
    "  --stats-output <path>        Write per-frame statistics and a summary to this file.\n"
    "  --camera <orbit|file>        Camera script: one frame per line, 9 numbers: forward xyz, up xyz,\n"
    "                               position xyz, in units where the volume spans (0,0,0)..(voxel ratio).\n"
    "                               The interactive 'orbit' camera needs a presentation backend.\n"
    "  --repeat <n>                 Render every camera frame n times.\n";
const char* HELP_CONVERT =
    "Usage: xenodon convert [options...] <source tiff path> <destination svo path>\n"
    "\n"
    "Options:\n"
    "  --dag              Merge identical subtrees (directed acyclic graph).\n"
    "  --rope             Store neighbour links in leaves, required by the svo-rope traversal.\n"
    "                     --dag and --rope are mutually exclusive.\n"
    "  --chan-diff <n>    Split a region while any channel differs by more than n (0..255). Default 0.\n"
    "  --std-dev <x>      Split a region while its standard deviation exceeds x (>= 0).\n"
    "                     --chan-diff and --std-dev are mutually exclusive.\n"
    "  --host             Build on the host even when a CUDA device is present (the GPU builder\n"
    "                     handles --chan-diff sparse/rope trees and writes identical files).\n";
const char* HELP_SYSINFO = "Usage: xenodon sysinfo\n\nLists CUDA devices; the index shown is the 'vkindex' of a headless configuration.\n";
const char* HELP_HEADLESS =
    "The headless configuration consists of one or more blocks:\n"
    "\n"
    "device {\n"
    "    vkindex = <gpu index>\n"
    "    offset = (<x>, <y>)\n"
    "    extent = (<width>, <height>)\n"
    "}\n"
    "\n"
    "Each block renders the rectangle offset/extent (pixels) of the total frame on the CUDA device\n"
    "<gpu index> (see 'xenodon sysinfo'). The same device may appear in several blocks. The frame is\n"
    "the union of all rectangles. Comments are not allowed.\n";

void help(const char* program, const std::vector<const char*>& args) {
    if (args.empty()) {
        std::fputs(HELP_MAIN, stdout);
        return;
    }
    if (args.size() != 1) {
        std::printf("Error: Invalid usage of subcommand 'help', see '%s help'\n", program);
        return;
    }
    const std::string_view topic = args[0];
    const std::pair<const char*, const char*> topics[] = {{"help", HELP_MAIN},       {"sysinfo", HELP_SYSINFO},
                                                           {"convert", HELP_CONVERT}, {"render", HELP_RENDER},
                                                           {"headless-config", HELP_HEADLESS}};
    for (const auto& t : topics)
        if (topic == t.first) {
            std::fputs(t.second, stdout);
            return;
        }
    std::printf("Error: No such topic %s, see '%s help'\n", args[0], program);
}

void sysinfo() {
    int n = 0;
    if (xn_device_count(&n) != XN_OK) {
        std::printf("Error: %s\n", xn_last_error());
        return;
    }
    std::printf("System setup information\nGPUs:\n");
    for (int i = 0; i < n; ++i) {
        char name[256] = "?";
        xn_device_name(i, name, sizeof name);
        std::printf("  GPU %d:\n    name: '%s'\n", i, name);
    }
    if (n == 0) std::printf("  (none)\n");
}

// ---------------------------------------------------------------------------------
// render
// ---------------------------------------------------------------------------------
struct RenderOptions {
    bool quiet = false, xorg = false, discard_output = false;
    std::string log_output, headless, output, direct, xorg_multi_gpu;
    std::string volume_path, volume_type, shader, stats_path, camera;
    float voxel_ratio[3] = {1, 1, 1};
    float emission = 1.0f;
    size_t repeat = 1;
};

Action voxel_ratio_opt(float* var) {
    return [var](const char* a) {
        const std::string_view arg = a;
        const size_t first = arg.find(':');
        const size_t second = first == std::string_view::npos ? first : arg.find(':', first + 1);
        if (first == std::string_view::npos || second == std::string_view::npos) return false;
        const std::string_view x = arg.substr(0, first), y = arg.substr(first + 1, second - first - 1),
                               z = arg.substr(second + 1);
        if (x.empty() || y.empty() || z.empty()) return false;
        float v[3];
        if (!parse_float(x, v[0]) || !parse_float(y, v[1]) || !parse_float(z, v[2])) return false;
        if (!(v[0] > 0 && v[1] > 0 && v[2] > 0)) return false;
        var[0] = v[0]; var[1] = v[1]; var[2] = v[2];
        return true;
    };
}

RenderOptions parse_render_args(const std::vector<const char*>& args) {
    RenderOptions o;
    ArgTable()
        .toggle("--quiet", 'q', &o.quiet)
        .toggle("--xorg", 0, &o.xorg)
        .toggle("--discard-output", 0, &o.discard_output)
        .option("--log-output", 0, "output path", string_opt(&o.log_output))
        .option("--headless", 0, "config path", string_opt(&o.headless))
        .option("--output", 0, "output path", string_opt(&o.output))
        .option("--direct", 0, "config path", string_opt(&o.direct))
        .option("--xorg-multi-gpu", 0, "config path", string_opt(&o.xorg_multi_gpu))
        .option("--emission-coeff", 'e', "emission coefficient", float_min_opt<float>(&o.emission, 0.f))
        .option("--volume-type", 0, "volume type", string_opt(&o.volume_type))
        .option("--shader", 's', "shader", string_opt(&o.shader))
        .option("--voxel-ratio", 'r', "voxel dimension ratio", voxel_ratio_opt(o.voxel_ratio))
        .option("--stats-output", 0, "stats output", string_opt(&o.stats_path))
        .option("--camera", 0, "camera", string_opt(&o.camera))
        .option("--repeat", 0, "frame repeat", int_range_opt<size_t>(&o.repeat, 0, std::numeric_limits<size_t>::max()))
        .operand("volume path", string_opt(&o.volume_path))
        .apply(args);

    const int backends = (int)o.xorg + (int)!o.headless.empty() + (int)!o.direct.empty();
    if (backends == 0) throw CliError("Missing required backend --xorg, --headless or --direct");
    if (backends > 1) throw CliError("--xorg, --headless and --direct are mutually exclusive");
    if (o.headless.empty() && o.discard_output) throw CliError("--dont-save requires --headless");
    if (!o.output.empty() && o.headless.empty()) throw CliError("--output requires --headless");
    else if (o.output.empty()) o.output = "out-{}.png";
    else if (o.discard_output) throw CliError("--dont-save and --output are mutually exclusive");
    if (!o.xorg_multi_gpu.empty() && !o.xorg) throw CliError("--xorg-multi-gpu requires --xorg");
    return o;
}

enum class FileType { Tiff, Svo, Unknown };
const char* file_type_name(FileType t) { return t == FileType::Tiff ? "tiff" : t == FileType::Svo ? "svo" : "unknown"; }
FileType parse_file_type(std::string_view s) {
    if (s == "tiff" || s == "tif") return FileType::Tiff;
    if (s == "svo") return FileType::Svo;
    return FileType::Unknown;
}
FileType guess_file_type(const RenderOptions& o) {
    if (!o.volume_type.empty()) return parse_file_type(o.volume_type);
    const size_t slash = o.volume_path.find_last_of('/');
    const std::string name = slash == std::string::npos ? o.volume_path : o.volume_path.substr(slash + 1);
    const size_t dot = name.find_last_of('.');
    if (dot == std::string::npos || dot == 0) return FileType::Unknown;
    return parse_file_type(std::string_view(name).substr(dot + 1));
}

struct ShaderOption {
    const char* option;
    FileType required;
    int traversal;
};
const ShaderOption SHADER_OPTIONS[] = {{"dda", FileType::Tiff, XN_DDA},        {"svo-naive", FileType::Svo, XN_SVO_NAIVE},
                                       {"esvo", FileType::Svo, XN_ESVO},       {"svo-df", FileType::Svo, XN_SVO_DF},
                                       {"svo-rope", FileType::Svo, XN_SVO_ROPE}};

const ShaderOption& select_shader(const RenderOptions& o, FileType type) {
    if (!o.shader.empty()) {
        for (const auto& s : SHADER_OPTIONS)
            if (o.shader == s.option) {
                if (s.required == type) return s;
                throw CliError("Shader '" + o.shader + "' is incompatible with model type '" + file_type_name(type) +
                               "' (requires '" + file_type_name(s.required) + "')");
            }
        throw CliError("Invalid shader '" + o.shader + "'");
    }
    for (const auto& s : SHADER_OPTIONS)
        if (s.required == type) return s;
    throw CliError("Failed to parse model file type");
}

// background PNG writer with a bounded queue
class FrameWriter {
    struct Job {
        std::string path;
        std::vector<uint32_t> pixels;
        uint32_t w, h;
    };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> queue;
    bool done = false;
    std::thread worker;

    void run() {
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lock(mu);
                cv.wait(lock, [&] { return done || !queue.empty(); });
                if (queue.empty()) return;
                job = std::move(queue.front());
            }
            LOGGER.log("Compressing...");
            if (xn_png_write(job.path.c_str(), job.pixels.data(), job.w, job.h) != XN_OK)
                LOGGER.log(std::string("Error saving output: ") + xn_last_error());
            else
                LOGGER.log("Saved output to '" + job.path + "'");
            {
                std::lock_guard<std::mutex> lock(mu);
                queue.pop_front();
            }
            cv.notify_all();
        }
    }

public:
    void submit(const std::string& path, const std::vector<uint32_t>& pixels, uint32_t w, uint32_t h) {
        if (!worker.joinable()) worker = std::thread([this] { run(); });
        std::unique_lock<std::mutex> lock(mu);
        cv.wait(lock, [&] { return queue.size() < 2; });
        queue.push_back(Job{path, pixels, w, h});
        cv.notify_all();
    }
    void finish() {
        {
            std::lock_guard<std::mutex> lock(mu);
            done = true;
        }
        cv.notify_all();
        if (worker.joinable()) worker.join();
    }
    ~FrameWriter() { finish(); }
};

struct Device {
    xn_ctx* ctx = nullptr;
    xn_rect region{};
};
struct Devices {
    std::vector<Device> v;
    ~Devices() {
        for (auto& d : v)
            if (d.ctx) xn_ctx_destroy(d.ctx);
    }
};

void main_loop(const RenderOptions& o, Devices& devs, const std::string& out_pattern) {
    // check_setup (src/main_loop.cpp:45-69): headless has one output per device
    {
        const size_t n = devs.v.size();
        std::string s = "Setup: " + std::to_string(n) + (n > 1 ? " devices" : " device") + ", with ";
        for (size_t i = 0; i < n; ++i) s += i == 0 ? "1" : ", 1";
        s += n > 1 ? " outputs" : " output";
        LOGGER.log(s);
    }

    const FileType type = guess_file_type(o);
    if (type == FileType::Unknown) throw CliError("Failed to parse model file type");
    LOGGER.log(std::string("Model file type: '") + file_type_name(type) + "'");
    const ShaderOption& shader = select_shader(o, type);
    LOGGER.log(std::string("Using shader '") + shader.option + "'");

    uint32_t model_dim[3];
    if (type == FileType::Tiff) {
        uint64_t dims[3];
        check(xn_tiff_info(o.volume_path.c_str(), dims));
        const uint64_t voxels = dims[0] * dims[1] * dims[2];
        std::printf("%llux%llux%llu = %llu pixels\n", (unsigned long long)dims[0], (unsigned long long)dims[1],
                    (unsigned long long)dims[2], (unsigned long long)voxels);
        // pipelined ingest per device: slices go from the file through page-locked staging to the
        // device, which decodes them (the volume is replicated; later devices read the page cache)
        for (auto& d : devs.v) check(xn_upload_grid_tiff(d.ctx, o.volume_path.c_str(), nullptr, nullptr));
        for (int i = 0; i < 3; ++i) model_dim[i] = (uint32_t)dims[i];
    } else {
        uint64_t side = 0, count = 0;
        check(xn_svo_info(o.volume_path.c_str(), &side, &count));
        std::vector<xn_node> nodes(count);
        check(xn_svo_read(o.volume_path.c_str(), nodes.data(), count));
        for (auto& d : devs.v) check(xn_upload_svo(d.ctx, nodes.data(), count, side));
        model_dim[0] = model_dim[1] = model_dim[2] = (uint32_t)side;
    }
    LOGGER.log("Model dimensions: " + std::to_string(model_dim[0]) + "x" + std::to_string(model_dim[1]) + "x" +
               std::to_string(model_dim[2]));

    // RenderContext::calculate_display_rect (src/render/RenderContext.cpp:43-60)
    xn_rect display = devs.v[0].region;
    for (size_t i = 1; i < devs.v.size(); ++i) {
        const xn_rect& r = devs.v[i].region;
        const int32_t x = std::min(display.x, r.x), y = std::min(display.y, r.y);
        const uint32_t w = std::max((uint32_t)display.x + display.w, (uint32_t)r.x + r.w) - (uint32_t)x;
        const uint32_t h = std::max((uint32_t)display.y + display.h, (uint32_t)r.y + r.h) - (uint32_t)y;
        display = xn_rect{x, y, w, h};
    }
    LOGGER.log("Total resolution: " + std::to_string(display.w) + "x" + std::to_string(display.h) + " pixels");
    for (auto& d : devs.v) {
        check(xn_set_target(d.ctx, &d.region, &display));
        check(xn_set_params(d.ctx, o.voxel_ratio, model_dim, o.emission));
    }

    // create_camera_controller (src/main_loop.cpp:176-189)
    if (o.camera.empty() || o.camera == "orbit")
        throw CliError("The orbit camera controller needs a presentation backend; pass --camera <file>");
    LOGGER.log("Reading camera transforms from '" + o.camera + "'");
    const std::string cam_text = read_text_file(o.camera, "Failed to open");
    int n_cam = 0;
    check(xn_camera_script_parse(cam_text.c_str(), nullptr, 0, &n_cam));
    std::vector<float> cams((size_t)n_cam * 9);
    check(xn_camera_script_parse(cam_text.c_str(), cams.data(), n_cam, &n_cam));

    std::vector<xn_ctx*> ctxs;
    for (auto& d : devs.v) ctxs.push_back(d.ctx);
    // Saved frames are compressed and written by a worker thread (two frames may be in flight),
    // so rendering frame i+1 overlaps the PNG encoding of frame i; the reference encodes
    // synchronously on its only thread (HeadlessDisplay.cpp:78-91).
    FrameWriter writer;
    std::vector<uint32_t> frame_pixels;
    if (!out_pattern.empty()) frame_pixels.resize((size_t)display.w * display.h);

    using clock = std::chrono::high_resolution_clock;
    auto start = clock::now();
    const auto run_start = start;
    size_t frames = 0, total_frames = 0, cam_index = 0;
    uint64_t saved_frame = 0;
    std::vector<xn_render_stats> all_stats;

    LOGGER.log("Starting render loop...");
    for (;;) {
        ++frames;
        ++total_frames;
        const float* cam = cams.data() + cam_index * 9;

        // MultiplexRenderer::render (src/render/MultiplexRenderer.cpp:21-31)
        for (auto& d : devs.v) check(xn_render(d.ctx, shader.traversal, cam, cam + 3, cam + 6));
        xn_render_stats st{0, 0, 0.0, 0.0, std::numeric_limits<double>::max()};
        for (auto& d : devs.v) {
            double ms = 0;
            check(xn_sync(d.ctx, &ms)); // HeadlessDisplay::swap_buffers: wait for every output
            st.total_rays += (uint64_t)d.region.w * d.region.h;
            st.outputs += 1;
            st.total_render_time += ms;
            st.max_render_time = std::max(st.max_render_time, ms);
            st.min_render_time = std::min(st.min_render_time, ms);
        }
        if (!out_pattern.empty()) {
            const std::string path = format_frame_name(out_pattern, saved_frame);
            LOGGER.log("Saving frame " + std::to_string(saved_frame) + "...");
            xn_rect enc;
            check(xn_frame_gather(ctxs.data(), (int)ctxs.size(), frame_pixels.data(), &enc));
            writer.submit(path, frame_pixels, enc.w, enc.h);
        }
        ++saved_frame;
        all_stats.push_back(st);

        // ScriptCameraController::update: the script's EOF ends the run
        if (o.repeat != 0 && total_frames % o.repeat == 0) {
            if (cam_index + 1 >= (size_t)n_cam) break;
            ++cam_index;
        }
        const auto now = clock::now();
        const std::chrono::duration<double> diff = now - start;
        if (diff > std::chrono::seconds{5}) {
            LOGGER.log("FPS: " + fmt_double((double)frames / diff.count()));
            frames = 0;
            start = now;
        }
    }
    writer.finish(); // all frames are on disk before the totals are reported
    const std::chrono::duration<double> total = clock::now() - run_start;

    uint64_t rays = 0;
    double ms = 0;
    for (const auto& s : all_stats) {
        rays += s.total_rays;
        ms += s.total_render_time;
    }
    LOGGER.log("total rays: " + std::to_string(rays) + ", total render time: " + fmt_double(ms) +
               "ms, mray/s: " + fmt_double((double)rays / (ms * 1000.0)));
    LOGGER.log("frames: " + std::to_string(all_stats.size()) + ", fps: " +
               fmt_double((double)all_stats.size() / total.count()) + ", total time: " + fmt_double(total.count()) + "s");
    if (!o.stats_path.empty()) {
        check(xn_stats_write(o.stats_path.c_str(), all_stats.data(), all_stats.size(), total.count()));
        LOGGER.log("Saved stats to '" + o.stats_path + "'");
    }
}

void render(const std::vector<const char*>& args) {
    RenderOptions o;
    try {
        o = parse_render_args(args);
    } catch (const CliError& e) {
        std::printf("Error: %s\n", e.what());
        return;
    }
    LOGGER.console = !o.quiet;
    if (!o.log_output.empty()) LOGGER.file.open(o.log_output);

    Devices devs;
    try {
        if (o.headless.empty())
            throw CliError("only the headless backend is available in this build (presentation backends are not part "
                           "of the traversal path)");
        const std::string conf = read_text_file(o.headless, "Failed to open config file");
        int n = 0;
        check(xn_headless_config_parse(conf.c_str(), nullptr, 0, &n));
        std::vector<xn_headless_device> entries((size_t)n);
        check(xn_headless_config_parse(conf.c_str(), entries.data(), n, &n));
        for (const auto& e : entries) {
            Device d;
            d.region = e.region;
            check(xn_ctx_create((int)e.vkindex, &d.ctx));
            devs.v.push_back(d);
        }
    } catch (const CliError& e) {
        std::printf("Error: Failed to initialize backend: %s\n", e.what());
        return;
    }

    try {
        main_loop(o, devs, o.discard_output ? std::string() : o.output);
    } catch (const CliError& e) {
        std::printf("Error: %s\n", e.what());
    }
}

// ---------------------------------------------------------------------------------
// convert (src/convert.cpp:13-122)
// ---------------------------------------------------------------------------------
void convert(const std::vector<const char*>& args) {
    std::string src, dst;
    bool dag = false, rope = false, host_only = false;
    int channel_difference = -1;
    double stddev = -1;
    ArgTable table;
    table.toggle("--dag", 0, &dag)
        .toggle("--rope", 0, &rope)
        .toggle("--host", 0, &host_only)
        .option("--chan-diff", 0, "channel difference", int_range_opt<int>(&channel_difference, 0, 255))
        .option("--std-dev", 0, "std. dev", float_min_opt<double>(&stddev, 0.0))
        .operand("source tiff path", string_opt(&src))
        .operand("destination svo path", string_opt(&dst));
    try {
        table.apply(args);
    } catch (const CliError& e) {
        std::printf("Error: %s\n", e.what());
        return;
    }
    if (dag && rope) {
        std::printf("Error: --dag and --rope are mutually exclusive\n");
        return;
    }
    if (channel_difference >= 0 && stddev >= 0) {
        std::printf("Error: --std-dev and --chan-diff are mutually exclusive\n");
        return;
    }

    std::printf("Loading source...\n");
    uint64_t dims[3];
    std::vector<uint8_t> grid;
    if (xn_tiff_info(src.c_str(), dims) != XN_OK) {
        std::printf("Error reading '%s': %s\n", src.c_str(), xn_last_error());
        return;
    }
    const uint64_t voxels = dims[0] * dims[1] * dims[2];
    std::printf("%llux%llux%llu = %llu pixels\n", (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)dims[2], (unsigned long long)voxels);
    // the reference's report (src/convert.cpp:66-71); memory_footprint() = sizeof(Grid) + voxels * 4,
    // sizeof(Grid) = sizeof(Octree) = 32 on x86-64 (src/model/Grid.h:62-64, Octree.h:66-68)
    constexpr unsigned long long OBJECT_BYTES = 32;
    std::printf("Source grid:\n Dimensions: %llux%llux%llu\n Size: %llu bytes\n", (unsigned long long)dims[0],
                (unsigned long long)dims[1], (unsigned long long)dims[2], OBJECT_BYTES + (unsigned long long)(voxels * 4));
    std::printf("Converting to octree...\n");
    std::fflush(stdout);
    xn_node* nodes = nullptr;
    uint64_t count = 0, side = 0;
    xn_build_stats st{};
    const int type = dag ? 1 : rope ? 2 : 0;
    // The GPU builder (byte-identical output) handles both heuristics and all three tree types; --host,
    // the absence of a CUDA device, or a --std-dev threshold within rounding distance of some cell's
    // deviation (the GPU builder refuses to guess) use the host builder.  On the GPU path the
    // volume goes from the file to the device through the ingest pipeline and never exists on the host.
    int rc = XN_ERR_INVALID;
    bool on_gpu = false;
    int n_devices = 0;
    if (!host_only && xn_device_count(&n_devices) == XN_OK && n_devices > 0) {
        xn_ctx* ctx = nullptr;
        if (xn_ctx_create(0, &ctx) == XN_OK) {
            xn_set_grid_layout(ctx, XN_GRID_LAYOUT_LINEAR); // the builder reads the x-major copy
            if (xn_upload_grid_tiff(ctx, src.c_str(), nullptr, nullptr) == XN_OK)
                rc = stddev >= 0 ? xn_convert_resident_grid_ex(ctx, 1, stddev, type, 0, &nodes, &count, &side, &st)
                                 : xn_convert_resident_grid(ctx, std::max(channel_difference, 0), type, 0, &nodes, &count, &side, &st);
            xn_ctx_destroy(ctx);
            on_gpu = rc == XN_OK;
        }
    }
    if (!on_gpu) {
        grid.resize(voxels * 4);
        if (xn_tiff_read(src.c_str(), grid.data(), grid.size()) != XN_OK) {
            std::printf("Error reading '%s': %s\n", src.c_str(), xn_last_error());
            return;
        }
        rc = stddev >= 0
                 ? xn_build_octree(grid.data(), dims[0], dims[1], dims[2], 1, stddev, type, &nodes, &count, &side, &st)
                 : xn_build_octree(grid.data(), dims[0], dims[1], dims[2], 0, (double)std::max(channel_difference, 0),
                                   type, &nodes, &count, &side, &st);
    }
    if (rc != XN_OK) {
        std::printf("Error: %s\n", xn_last_error());
        return;
    }
    std::fprintf(stderr, "Built on the %s\n", on_gpu ? "GPU" : "host"); // not part of the reference's report
    {
        // same arithmetic as the reference's report (src/convert.cpp:86-111), quirks included
        auto ipow = [](size_t x, size_t y) {
            size_t z = 1;
            while (--y) z *= x;
            return z;
        };
        const size_t perfect = (ipow(8, st.depth + 1) - 1) / (8 + 1);
        const double total_prop = (double)st.total_nodes / (double)perfect;
        const double unique_prop = (double)count / (double)perfect;
        std::printf("Generated octree:\n");
        std::printf(" Dimensions: %llux%llux%llu\n", (unsigned long long)side, (unsigned long long)side,
                    (unsigned long long)side);
        std::printf(" Size: %llu bytes\n", OBJECT_BYTES + (unsigned long long)(count * sizeof(xn_node)));
        std::printf(" Perfect tree nodes: %llu\n", (unsigned long long)perfect);
        std::printf(" Total nodes: %llu (%.5f%%)\n", (unsigned long long)st.total_nodes, total_prop * 100);
        std::printf(" Unique nodes: %llu (%.5f%%)\n", (unsigned long long)count, unique_prop * 100);
        std::printf(" Total leaves: %llu\n", (unsigned long long)st.total_leaves);
        std::printf(" Unique leaves: %llu\n", (unsigned long long)st.unique_leaves);
        std::printf(" Depth: %llu\n", (unsigned long long)st.depth);
    }
    if (xn_svo_write(dst.c_str(), nodes, count, side) != XN_OK)
        std::printf("Error writing '%s': %s\n", dst.c_str(), xn_last_error());
    xn_free(nodes);
}

} // namespace

int main(int argc, const char* argv[]) {
    if (argc < 2) {
        help(argv[0], {});
        return 0;
    }
    const std::vector<const char*> args(argv + 2, argv + argc);
    const std::string_view sub = argv[1];
    if (sub == "help") help(argv[0], args);
    else if (sub == "sysinfo") sysinfo();
    else if (sub == "render") render(args);
    else if (sub == "convert") convert(args);
    else std::printf("Error: Invalid subcommand '%s', see '%s help'\n", argv[1], argv[0]);
    return EXIT_SUCCESS; // the reference always exits 0 (src/main.cpp:245)
}
