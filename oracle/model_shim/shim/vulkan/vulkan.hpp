// Stand-in for <vulkan/vulkan.hpp>: the reference's model code (src/model, src/utility/Span.h)
// only needs the *type* vk::ArrayProxy to exist.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
namespace vk {
template <typename T>
struct ArrayProxy {
    size_t n = 0;
    const T* p = nullptr;
    size_t size() const { return n; }
    const T* data() const { return p; }
};
} // namespace vk
