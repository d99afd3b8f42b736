// vkrt_internal.h -- host-side declarations shared by the translation units of libvkrt_b200.
#pragma once
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/vkrt_b200.h"

namespace vkrt {

struct InstanceXform;
struct SceneView;
struct Wavefront;
struct TraceStats;

struct BvhResult
{
  float4*  nodes    = nullptr;
  float4*  tris     = nullptr;  // inside the nodes allocation
  size_t   arenaBytes = 0;
  uint2*   triInfo  = nullptr;
  uint4*   triShade = nullptr;  // 7 x uint4 per flattened triangle
  uint32_t numNodes = 0;
  uint32_t numTris  = 0;
  int      maxDepth = 0;
  // 32-byte quantised copy of `nodes` (the format the trace kernels read by default), see quantizeBvh
  uint4*   qnodes   = nullptr;  // 2 per node
  float4*  cnodes   = nullptr;  // 4 per node: centre / half-extent form of `nodes` (default traversal format)
  uint4*   wnodes   = nullptr;  // 4 per BVH4 node (64 bytes), see collapseBvhWide
  uint32_t numWideNodes = 0;
  int      wideDepth = 0;
  float    gridOrigin[3] = {0.f, 0.f, 0.f};
  float    gridInv[3]    = {1.f, 1.f, 1.f};
};

// vkrt_bvh.cu
bool buildBvh(cudaStream_t stream, const InstanceXform* dInst, const uint32_t* dInstTriOffset, uint32_t numInst, const InstanceData* dGeo,
              uint32_t numTris, int buildMode, BvhResult& out, std::string& err, uint64_t* launches);

// 64-byte fp32 nodes -> 32-byte nodes: child boxes on a 16-bit grid over the root box (conservatively padded), same topology
bool quantizeBvh(cudaStream_t stream, BvhResult& bvh, std::string& err, uint64_t* launches);
// binary tree -> 4-wide nodes on the same grid (call after quantizeBvh)
bool collapseBvhWide(cudaStream_t stream, BvhResult& bvh, std::string& err, uint64_t* launches);

// vkrt_sort.cu: optional ordering of the ray queue between bounces (octant + origin Morton buckets), hand written, device side only
size_t sortScratchBytes(uint32_t maxItems);
void   launchSortQueue(cudaStream_t s, const SceneView& scene, const float4* rayO, const float4* rayD, const uint32_t* queue, uint32_t* queueOut,
                       const uint32_t* count, uint32_t capacity, void* scratch, int numSMs);

// vkrt_trace.cu
struct TraceLaunch
{
  const float4*   rayO;
  const float4*   rayD;
  const uint32_t* queue;      // ray slots to trace (nullptr: identity)
  const uint32_t* count;      // device count of queue entries
  uint32_t*       cursor;     // device work cursor (zeroed by the caller)
  float4*         hit;        // closest: per ray slot {t,u,v,tri}
  // shadow resolve (any-hit): adds contribution to rad[path] when unoccluded
  const float4*   shC;
  float4*         rad;
  TraceStats*     stats;      // nullptr unless counting
  uint32_t        primary;    // shadow resolve: this is bounce 0 (bounce-class statistics)
  float4*         cost;       // heat-map debug mode: per path, .w accumulates the traversal steps of its rays (counting kernels only)
};
int  traceBvhWidth(uint32_t numTris);  // 2 or 4: tree the trace kernels walk for a scene of this size (VKRT_BVH_WIDTH)
int  traceNodeBytes(int bvhWidth);     // bytes per node the trace kernels fetch: 64, or 32 with VKRT_NODE_FORMAT=q32
void launchTraceClosest(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, bool count, int numSMs);
void launchTraceShadow(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, bool count, int numSMs);
void launchTraceAnyProbe(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs);

// vkrt_shade.cu
struct FrameParams
{
  RtxState state;       // frame = first frame of the batch
  int32_t  seedMode;
  int32_t  framesInBatch;
  int32_t  sampleIndex; // 0..maxSamples-1
  int32_t  depth;
};
void launchRaygen(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, int numSMs);
// vkrt_shade.cu is compiled twice: IEEE-exact arithmetic, and approximate division / sqrt / SFU transcendentals (the default)
void launchShadePrecise(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs);
void launchShadeFast(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs);
// fused tail (vkrt_shade.cu): every remaining bounce from fp.depth on, one thread per live path
void launchTailPrecise(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs);
void launchTailFast(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs);
void launchFinishSample(cudaStream_t s, const Wavefront& wf, const FrameParams& fp, int numSMs);
// fromRadiance: maxSamples == 1 -- the firefly clamp of finishSample is applied here and that launch is skipped
// mirror: optional second image (a peer's, over NVLink) that receives the same stores
void launchAccumulate(cudaStream_t s, const Wavefront& wf, const FrameParams& fp, float4* image, float4* mirror, int numSMs, bool fromRadiance);
void launchShardCopy(cudaStream_t s, float4* image, const uint32_t* pixels, uint32_t n, float4* compact, bool toImage);
struct MipView  // RGBA32F mip pyramid of the accumulation image (level 0 = the image itself)
{
  const float4* level[16];
  int           w[16], h[16];
  int           count;
};
void launchMipDown(cudaStream_t s, const float4* src, int sw, int sh, float4* dst, int dw, int dh);
void launchTonemap(cudaStream_t s, const float4* image, int w, int h, const Tonemapper& tm, const MipView& mips, float4* outF, uchar4* out8);

// vkrt_host.cpp
void buildEnvAccelHost(const float* rgba, uint32_t w, uint32_t h, EnvAccel* out, float* integral, float* average);
bool invertAffine(const float* m16, float* inv16);
bool buildSahHost(const float* tris, uint32_t n, std::vector<float>& nodesOut, std::vector<uint32_t>& orderOut, int& maxDepthOut);

}  // namespace vkrt
