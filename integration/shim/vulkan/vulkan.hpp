// Stand-in for <vulkan/vulkan.hpp>, written from the names the reference's headers mention, so that
// integration/CudaMultiplexRenderer.h can be COMPILED against the reference's own headers
// (render/RenderStats.h, backend/headless/HeadlessConfig.h, model/*.h ...) in an image without the
// Vulkan SDK.  Types are empty shells with just the members those headers touch inline; nothing
// here is ever executed.  TEST INFRASTRUCTURE ONLY (tests/test_integration_adapter.py).
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
template <typename T>
struct UniqueHandle {
    T h{};
    const T& get() const { return h; }
    T& get() { return h; }
    const T& operator*() const { return h; }
    T& operator*() { return h; }
    const T* operator->() const { return &h; }
    T* operator->() { return &h; }
};
struct Offset2D { int32_t x = 0, y = 0; };
struct Extent2D { uint32_t width = 0, height = 0; };
struct Rect2D { Offset2D offset; Extent2D extent; };
enum class Result { eSuccess };
enum class PhysicalDeviceType { eOther };
enum class PresentModeKHR { eFifoKHR };
enum class CommandBufferUsageFlagBits { eOneTimeSubmit };
enum class PipelineStageFlagBits { eColorAttachmentOutput };
struct PipelineStageFlags { PipelineStageFlags(PipelineStageFlagBits = {}) {} };
struct QueueFlags {};
struct MemoryPropertyFlags {};
struct MemoryRequirements {};
struct AttachmentDescription {};
struct DeviceQueueCreateInfo {};
struct SurfaceFormatKHR {};
struct PhysicalDeviceProperties { PhysicalDeviceType deviceType{}; char deviceName[256]{}; };
struct CommandBufferBeginInfo { CommandBufferBeginInfo(CommandBufferUsageFlagBits = {}) {} };
struct CommandBuffer {
    void begin(const CommandBufferBeginInfo&) const {}
    void end() const {}
};
struct SubmitInfo { uint32_t commandBufferCount = 0; const CommandBuffer* pCommandBuffers = nullptr; };
struct Fence {};
struct Queue {
    void submit(const SubmitInfo&, Fence) const {}
    void waitIdle() const {}
};
struct Device { Queue getQueue(uint32_t, uint32_t) const { return {}; } };
struct CommandPool {};
struct DeviceMemory {};
struct DisplayKHR {};
struct Image {};
struct ImageView {};
struct PhysicalDevice {};
struct Semaphore {};
struct SurfaceKHR {};
struct QueryPool {};
struct SwapchainKHR {};
using UniqueCommandBuffer = UniqueHandle<CommandBuffer>;
using UniqueDevice = UniqueHandle<Device>;
using UniqueDeviceMemory = UniqueHandle<DeviceMemory>;
using UniqueFence = UniqueHandle<Fence>;
using UniqueImageView = UniqueHandle<ImageView>;
using UniqueQueryPool = UniqueHandle<QueryPool>;
using UniqueSemaphore = UniqueHandle<Semaphore>;
using UniqueSwapchainKHR = UniqueHandle<SwapchainKHR>;
} // namespace vk
