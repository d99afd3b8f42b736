// vkrt_sort.cu -- optional ray sort between bounces (vkrt_render_opts.sort_rays).
//
// The shade kernel already compacts the live paths of a bounce into a dense queue (warp-ballot aggregated appends,
// vkrt_shade.cu).  This pass additionally orders that queue so that neighbouring lanes carry similar rays: the key is the
// direction octant (3 bits) followed by a 27-bit Morton code of the ray origin on the BVH's 16-bit scene grid (9 bits per
// axis).  Traversal and shading results do not depend on the order of the queue (every path owns its state and its RNG
// stream), so the image is bit-identical with and without the sort; it only changes how coherent a warp's memory accesses are.
// The reference has no counterpart: a megakernel keeps each path in its invocation (shaders/pathtrace.glsl:193-343).
#include <cub/cub.cuh>
#include "vkrt_device.cuh"
#include "vkrt_internal.h"

namespace vkrt {

__device__ __forceinline__ uint32_t spread9(uint32_t v)  // 9 bits -> every third bit
{
  v &= 0x1ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) sortKeyKernel(const __grid_constant__ SceneView sc, const float4* __restrict__ rayO,
                                                     const float4* __restrict__ rayD, const uint32_t* __restrict__ queue, uint32_t count,
                                                     uint32_t* __restrict__ keys)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= count)
    return;
  const uint32_t p = queue[i];
  const float4   o = rayO[p], d = rayD[p];
  const float gx = fminf(fmaxf((o.x - sc.gridOrigin.x) * sc.gridInv.x, 0.0f), 65535.0f);
  const float gy = fminf(fmaxf((o.y - sc.gridOrigin.y) * sc.gridInv.y, 0.0f), 65535.0f);
  const float gz = fminf(fmaxf((o.z - sc.gridOrigin.z) * sc.gridInv.z, 0.0f), 65535.0f);
  const uint32_t m = (spread9((uint32_t)gx >> 7) << 2) | (spread9((uint32_t)gy >> 7) << 1) | spread9((uint32_t)gz >> 7);
  const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
  keys[i] = (oct << 27) | m;
}

size_t sortTempBytes(uint32_t maxItems)
{
  size_t bytes = 0;
  cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, (int)maxItems, 0, 30);
  return bytes;
}

// Sorts queue[0..count) by ray key.  keysA/keysB and queue/queueAlt are ping-pong buffers; returns which value buffer holds the
// sorted queue (0: queue, 1: queueAlt).
int launchSortQueue(cudaStream_t s, const SceneView& scene, const float4* rayO, const float4* rayD, uint32_t* queue, uint32_t* queueAlt,
                    uint32_t count, uint32_t* keysA, uint32_t* keysB, void* temp, size_t tempBytes)
{
  sortKeyKernel<<<(count + 255) / 256, 256, 0, s>>>(scene, rayO, rayD, queue, count, keysA);
  cub::DoubleBuffer<uint32_t> k(keysA, keysB), v(queue, queueAlt);
  cub::DeviceRadixSort::SortPairs(temp, tempBytes, k, v, (int)count, 0, 30, s);
  return v.Current() == queue ? 0 : 1;
}

}  // namespace vkrt
