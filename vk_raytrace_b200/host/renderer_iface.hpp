// renderer_iface.hpp -- the reference's abstract back-end interface (src/renderer.h:30-48), restated so that the
// adaptor below compiles with or without the Vulkan SDK / nvpro_core:
//   * inside the reference tree, define VKRT_HAVE_NVPRO: the real "renderer.h" (and through it the reference's
//     shaders/host_device.h) is included instead, and include/vkrt_host_device.h then only asserts the struct sizes
//     (tests/test_abi_host.py compiles the adaptor that way against /root/reference/src/renderer.h);
//   * standalone (this repository, no Vulkan SDK), the Vulkan / nvvk types are opaque stand-ins of the right kind
//     (dispatchable handles are pointers, VkExtent2D is {width,height}) and the class below is the reference's declaration
//     member for member -- no virtual destructor there, none here -- so signatures and vtable shape match.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#ifdef VKRT_HAVE_NVPRO
#include "renderer.h"  // src/renderer.h of the reference tree (pulls nvvk + shaders/host_device.h)
#endif
#include "../../include/vkrt_b200.h"

#ifndef VKRT_HAVE_NVPRO
typedef struct VkDevice_T*              VkDevice;
typedef struct VkPhysicalDevice_T*      VkPhysicalDevice;
typedef struct VkCommandBuffer_T*       VkCommandBuffer;
typedef struct VkDescriptorSet_T*       VkDescriptorSet;
typedef struct VkDescriptorSetLayout_T* VkDescriptorSetLayout;
struct VkExtent2D
{
  uint32_t width, height;
};
namespace nvvk {
class ResourceAllocator;
class ProfilerVK;
}  // namespace nvvk
class Scene;

class Renderer
{
public:
  virtual void setup(const VkDevice& device, const VkPhysicalDevice& physicalDevice, uint32_t familyIndex, nvvk::ResourceAllocator* allocator) = 0;
  virtual void destroy() = 0;
  virtual void run(const VkCommandBuffer& cmdBuf, const VkExtent2D& size, nvvk::ProfilerVK& profiler,
                   const std::vector<VkDescriptorSet>& extraDescSets) = 0;
  virtual void create(const VkExtent2D& size, const std::vector<VkDescriptorSetLayout>& extraDescSetsLayout, Scene* _scene = nullptr) = 0;
  virtual const std::string name() = 0;
  void                      setPushContants(const RtxState& state) { m_state = state; }

  RtxState m_state{};
};
#endif
