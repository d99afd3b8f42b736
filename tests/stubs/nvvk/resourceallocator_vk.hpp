// TEST STUB (tests/test_abi_host.py::test_adaptor_compiles_against_reference_header): the least of nvpro_core / Vulkan /
// glm that /root/reference/src/renderer.h and shaders/host_device.h need in order to be *compiled*: opaque handle types and
// packed glm-like vectors.  Nothing here is product code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
typedef struct VkDevice_T*              VkDevice;
typedef struct VkPhysicalDevice_T*      VkPhysicalDevice;
typedef struct VkCommandBuffer_T*       VkCommandBuffer;
typedef struct VkDescriptorSet_T*       VkDescriptorSet;
typedef struct VkDescriptorSetLayout_T* VkDescriptorSetLayout;
struct VkExtent2D
{
  uint32_t width, height;
};
namespace glm {
template <typename T, int N>
struct vecN
{
  T        v[N];
  T&       operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};
using vec2  = vecN<float, 2>;
using vec3  = vecN<float, 3>;
using vec4  = vecN<float, 4>;
using ivec2 = vecN<int, 2>;
struct mat4
{
  vec4        c[4];
  vec4&       operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
}  // namespace glm
namespace nvvk {
class ResourceAllocator;
}
