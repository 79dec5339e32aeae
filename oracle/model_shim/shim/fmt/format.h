// Stand-in for fmt 5.3 <fmt/format.h> (not vendored by the reference, subprojects/fmt.wrap).
// The model code only formats log/exception text with it; no arithmetic depends on it, so
// formatting is reduced to "return the pattern".  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
#include <string_view>
namespace fmt {
struct format_args {};
template <typename... A>
format_args make_format_args(const A&...) { return {}; }
template <typename... A>
std::string format(std::string_view f, const A&...) { return std::string(f); }
template <typename... A>
void print(std::string_view, const A&...) {}
} // namespace fmt
