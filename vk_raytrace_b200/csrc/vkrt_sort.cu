// vkrt_sort.cu -- optional ray sort between bounces (vkrt_render_opts.sort_rays): a hand-written bucket sort that runs entirely
// on the device (no library sort, no host synchronisation: the live count is read from the wavefront counters by the kernels).
//
// The shade kernel already compacts the live paths of a bounce into a dense queue (warp-ballot aggregated appends,
// vkrt_shade.cu).  This pass additionally orders that queue so that neighbouring lanes carry similar rays.  Key = direction
// octant (3 bits) + Morton code of the ray origin on a 16^3 grid over the scene (12 bits): 32768 buckets.
//   1. sortCountKernel    key per live path + bucket histogram (global atomics, aggregated per warp with match.any)
//   2. sortScanKernel     one block: exclusive prefix sum of the 32768 bucket counts (each thread 32 buckets + block scan)
//   3. sortScatterKernel  each path takes the next slot of its bucket (warp-aggregated atomic) in the output queue
// Order inside a bucket is arrival order, which varies from run to run -- and does not matter: traversal and shading results do not
// depend on the order of the queue (every path owns its state and its RNG stream), so the image is bit-identical with and without
// the sort (tests/test_gpu_parity.py).  It only changes how coherent a warp's memory accesses are.
// The reference has no counterpart: a megakernel keeps each path in its invocation (shaders/pathtrace.glsl:193-343).
#include <algorithm>
#include "vkrt_device.cuh"
#include "vkrt_internal.h"

namespace vkrt {

constexpr uint32_t kSortBuckets = 1u << 15;

__device__ __forceinline__ uint32_t spread4(uint32_t v)  // 4 bits -> every third bit
{
  v &= 0xfu;
  v = (v | (v << 4)) & 0xc3u;
  v = (v | (v << 2)) & 0x249u;
  return v;
}

__device__ __forceinline__ uint32_t rayBucket(const SceneView& sc, const float4 o, const float4 d)
{
  const float    gx = fminf(fmaxf((o.x - sc.gridOrigin.x) * sc.gridInv.x, 0.0f), 65535.0f);
  const float    gy = fminf(fmaxf((o.y - sc.gridOrigin.y) * sc.gridInv.y, 0.0f), 65535.0f);
  const float    gz = fminf(fmaxf((o.z - sc.gridOrigin.z) * sc.gridInv.z, 0.0f), 65535.0f);
  const uint32_t m  = (spread4((uint32_t)gx >> 12) << 2) | (spread4((uint32_t)gy >> 12) << 1) | spread4((uint32_t)gz >> 12);
  const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
  return (oct << 12) | m;
}

__global__ void __launch_bounds__(256) sortCountKernel(const __grid_constant__ SceneView sc, const float4* __restrict__ rayO, const float4* __restrict__ rayD,
                                                       const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count,
                                                       uint32_t* __restrict__ keys, uint32_t* __restrict__ hist)
{
  const uint32_t n = *count;
  for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const uint32_t p = queue[i];
    const uint32_t k = rayBucket(sc, rayO[p], rayD[p]);
    keys[i]          = k;
    // lanes of a warp with the same bucket (neighbouring pixels: common) share one atomic
    const unsigned m = __match_any_sync(__activemask(), k);
    if((int)(threadIdx.x & 31u) == __ffs(m) - 1)
      atomicAdd(&hist[k], (uint32_t)__popc(m));
  }
}

// exclusive scan of hist[0..kSortBuckets) in place -> bucket start offsets (1024 threads x 32 buckets)
__global__ void __launch_bounds__(1024) sortScanKernel(uint32_t* __restrict__ hist)
{
  __shared__ uint32_t warpSum[32];
  const uint32_t t = threadIdx.x, base = t * (kSortBuckets / 1024);
  uint32_t       local[kSortBuckets / 1024];
  uint32_t       sum = 0;
#pragma unroll
  for(uint32_t k = 0; k < kSortBuckets / 1024; k++)
  {
    local[k] = sum;
    sum += hist[base + k];
  }
  // block exclusive scan of `sum`
  uint32_t incl = sum;
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
  {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if((int)(t & 31u) >= o) incl += v;
  }
  if((t & 31u) == 31u) warpSum[t >> 5] = incl;
  __syncthreads();
  if(t < 32)
  {
    uint32_t w = warpSum[t], wi = w;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1)
    {
      const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
      if((int)t >= o) wi += v;
    }
    warpSum[t] = wi - w;  // exclusive
  }
  __syncthreads();
  const uint32_t excl = warpSum[t >> 5] + incl - sum;
#pragma unroll
  for(uint32_t k = 0; k < kSortBuckets / 1024; k++)
    hist[base + k] = excl + local[k];
}

__global__ void __launch_bounds__(256) sortScatterKernel(const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count, const uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ offsets, uint32_t* __restrict__ queueOut)
{
  const uint32_t n = *count;
  for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const uint32_t k    = keys[i];
    const unsigned m    = __match_any_sync(__activemask(), k);
    const int      lead = __ffs(m) - 1;
    uint32_t       pos  = 0;
    if((int)(threadIdx.x & 31u) == lead)
      pos = atomicAdd(&offsets[k], (uint32_t)__popc(m));
    pos = __shfl_sync(m, pos, lead) + (uint32_t)__popc(m & ((1u << (threadIdx.x & 31u)) - 1u));
    queueOut[pos] = queue[i];
  }
}

size_t sortScratchBytes(uint32_t maxItems) { return 4ull * maxItems + 4ull * kSortBuckets; }

// Orders queue[0..*count) by ray bucket into queueOut.  scratch: sortScratchBytes(capacity) bytes (keys, then the histogram).
void launchSortQueue(cudaStream_t s, const SceneView& scene, const float4* rayO, const float4* rayD, const uint32_t* queue, uint32_t* queueOut,
                     const uint32_t* count, uint32_t capacity, void* scratch, int numSMs)
{
  uint32_t* keys = (uint32_t*)scratch;
  uint32_t* hist = keys + capacity;
  cudaMemsetAsync(hist, 0, 4ull * kSortBuckets, s);
  const int grid = std::max(1, std::min<int>((int)((capacity + 255u) / 256u), numSMs * 8));
  sortCountKernel<<<grid, 256, 0, s>>>(scene, rayO, rayD, queue, count, keys, hist);
  sortScanKernel<<<1, 1024, 0, s>>>(hist);
  sortScatterKernel<<<grid, 256, 0, s>>>(queue, count, keys, hist, queueOut);
}

}  // namespace vkrt
