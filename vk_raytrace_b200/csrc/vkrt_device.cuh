// vkrt_device.cuh -- device-side views of the scene, the BVH and the wavefront state.
//
// HBM layout (DESIGN.md "data layout"):
//   BVH2 node     64 B = 4 x float4  {c0.lo.x c0.hi.x c0.lo.y c0.hi.y | c1.lo.x c1.hi.x c1.lo.y c1.hi.y |
//                                      c0.lo.z c0.hi.z c1.lo.z c1.hi.z | child0 child1 - -}
//   triangle      48 B = 3 x float4  {v0.xyz flags | v1.xyz triId | v2.xyz -}   (world space, BVH leaf order)
//   ray           32 B = 2 x float4  {o.xyz tmax | d.xyz seed}
//   hit           16 B = 1 x float4  {t u v triId}
//   path state    48 B = 3 x float4  {throughput.xyz seed | radiance.xyz flags | absorption.xyz sampleSum-slot}
#pragma once
#include "vkrt_math.cuh"
#include "../../include/vkrt_host_device.h"

namespace vkrt {

enum TriFlags : uint32_t { TRI_OPAQUE = 1u, TRI_CULL_DISABLE = 2u, TRI_FLIP_FACING = 4u };

// child encoding in BVH nodes
constexpr uint32_t kLeafBit   = 0x80000000u;
constexpr uint32_t kLeafCntSh = 28;           // bits 28..30: triangle count - 1
constexpr uint32_t kLeafMask  = 0x0fffffffu;  // bits 0..27 : first triangle slot

struct InstanceXform  // one per TLAS-equivalent node (src/accelstruct.cpp:132-157), 112 B
{
  float    o2w[12];  // row-major 3x4
  float    w2o[12];
  uint32_t primMesh;
  uint32_t flags;
  uint32_t triOffset;
  uint32_t triCount;
};

struct TexDesc
{
  const uchar4* data;
  int32_t       w, h;
  int32_t       linear;  // mag filter (LOD 0)
  int32_t       wrapS, wrapT;  // 0 repeat 1 clamp 2 mirror
  int32_t       _pad;
};

// What the any-hit alpha test needs from a material, in one place (80 bytes, by material index): the traversal kernels
// reach it from the triangle record alone, in parallel with the triangle's texture coordinates, instead of walking
// triangle -> instance -> geometry -> material -> texture table.
struct AlphaMat
{
  float   alphaFactor;  // pbrBaseColorFactor.a
  float   alphaCutoff;
  int32_t alphaMode;
  int32_t hasTexture;   // pbrBaseColorTexture > -1
  float   uvRows[8];    // the two rows of uvTransform that applyUvTransform reads
  TexDesc texture;
};

struct SceneView
{
  // traversal
  const float4* nodes;     // 4 per node (fp32 boxes, 64 bytes)
  const float4* cnodes;    // 4 per node (centre / half-extent boxes, 64 bytes): the default traversal format
  const uint4*  qnodes;    // 2 per node (16-bit grid boxes, 32 bytes): what the trace kernels read by default
  const uint4*  wnodes;    // 4 per BVH4 node (four 16-bit grid boxes + four references, 64 bytes), VKRT_BVH_WIDTH=4
  float3        gridOrigin, gridInv;  // G(p) = (p - gridOrigin) * gridInv, see quantizeBvh
  const float4* tris;      // 3 per triangle, leaf order
  const uint2*  triInfo;   // by flattened triangle id: {node(instance), prim}
  const uint4*  triShade;  // by flattened triangle id: 3 x VertexAttributes (6 x uint4) + {instance, material, flags, prim}
  // shading
  const InstanceXform*     instances;
  const InstanceData*      geoInfo;    // per primMesh: reference layout, CUDA pointers
  const GltfShadeMaterial* materials;
  const AlphaMat*          alphaMats;  // by material index
  const Light*             lights;
  const TexDesc*           textures;
  const float4*            envPixels;
  const EnvAccel*          envAccel;
  int32_t                  envW, envH;
  uint32_t                 numTris;
  uint32_t                 bvhWidth;   // 2 or 4: which tree the trace kernels walk for this scene
  uint32_t                 hasNonOpaque;
  SceneCamera              camera;
  SunAndSky                sunsky;
};

// Wavefront buffers (SoA, one entry per path in flight).
struct Wavefront
{
  float4*   rayO;      // o.xyz, tmax
  float4*   rayD;      // d.xyz, seed bits (seed at the time of the trace call)
  float4*   hit;       // t,u,v,tri
  float4*   thr;       // throughput.xyz, seed bits
  float4*   rad;       // radiance.xyz, flags
  float4*   absb;      // absorption.xyz, -
  float4*   sum;       // per-pixel sample sum for maxSamples > 1 (xyz), -
  float4*   shO;       // shadow rays: o.xyz, tmax
  float4*   shD;       // d.xyz, seed bits
  float4*   shC;       // contribution.xyz, path id bits
  uint32_t* queueCur;  // path ids to process this bounce
  uint32_t* queueNext;
  // Per bounce d four counters at counters[4 d ..]: [0] paths queued for the closest-hit trace of bounce d (raygen / shade of bounce
  // d - 1 fill it), [1] its work cursor, [2] shadow rays of bounce d, [3] their cursor.  Raygen zeroes all of them, so nothing has to
  // be reset or swapped between bounces (round 1 launched a one-thread kernel per bounce for that).
  uint32_t* counters;
  uint32_t  cbase;     // 4 x the bounce this launch works on
  uint32_t  clevels;   // bounces the counter array holds (maxDepth + 1)
  const uint32_t* pixelOfLocal;  // local pixel slot -> y*W + x
  uint32_t  numLocalPixels;
  uint32_t  numPaths;  // numLocalPixels * framesInFlight
};

// Wavefront state streams through the GPU once per bounce (gigabytes per launch) while the scene -- BVH, shading records, materials,
// textures: ~100 MB -- is reused by every ray.  VKRT_STREAM_HINTS marks the state accesses evict-first (ld.global.cs / st.global.cs) so
// that they do not push the scene out of the 126 MB L2.
#ifndef VKRT_STREAM_HINTS
#define VKRT_STREAM_HINTS 1
#endif
template <class T>
VKRT_D T ldS(const T* p)
{
#if VKRT_STREAM_HINTS
  return __ldcs(p);
#else
  return *p;
#endif
}
template <class T>
VKRT_D void stS(T* p, T v)
{
#if VKRT_STREAM_HINTS
  __stcs(p, v);
#else
  *p = v;
#endif
}

// nanosecond timer of the device (PTX %globaltimer, low word): what the heat-map debug view measures per path, like the
// reference's clockRealtimeEXT around its per-pixel loop (shaders/pathtrace.comp:89,108-117)
VKRT_D uint32_t deviceNs()
{
  uint32_t t;
  asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t));
  return t;
}

struct TraceStats  // device counters
{
  unsigned long long closestRays, shadowRays, nodesVisited, trisTested, shadowNodes, shadowTris, shadowRaysPrimary;
};

}  // namespace vkrt
