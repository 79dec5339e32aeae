// Stand-in for fmt 5.3 <fmt/format.h> (the reference downloads it through subprojects/fmt.wrap; it
// is not vendored).  Enough surface for the reference's headers to compile; formatting returns the
// pattern.  TEST INFRASTRUCTURE ONLY (tests/test_integration_adapter.py).
#pragma once
#include <string>
#include <string_view>
namespace fmt {
struct format_args {};
struct memory_buffer { std::string s; };
template <typename... A>
format_args make_format_args(const A&...) { return {}; }
template <typename... A>
std::string format(std::string_view f, const A&...) { return std::string(f); }
template <typename... A>
void print(std::string_view, const A&...) {}
template <typename... A>
void format_to(memory_buffer& b, std::string_view f, const A&...) { b.s += f; }
inline void vformat_to(memory_buffer& b, std::string_view f, format_args) { b.s += f; }
inline std::string to_string(const memory_buffer& b) { return b.s; }
} // namespace fmt
