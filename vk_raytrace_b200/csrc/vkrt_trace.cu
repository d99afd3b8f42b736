// vkrt_trace.cu -- persistent-thread BVH2 traversal kernels for sm_100a.
//
// Replaces the driver traversal behind ClosestHit()/AnyHit() of the reference
// (shaders/traceray_rq.glsl:108-185, shaders/traceray_rtx.glsl:30-73) and the inline
// alpha test HitTest()/pathtrace.rahit (traceray_rq.glsl:32-102).
//
// Kernel shape (B200 has no RT cores: traversal is SM + L1/L2 work, so SIMT lane
// utilisation is the first-order cost):
//   * grid = SMs x resident CTAs, persistent; every lane owns one ray at a time and
//     idle lanes are refilled from a global cursor as soon as fewer than
//     kRefetchThreshold lanes of the warp are still traversing (dynamic fetch);
//   * "while-while" loop with a postponed leaf: a lane that reaches a leaf parks it and
//     keeps descending until every lane of the warp has a leaf, then the warp
//     intersects triangles together;
//   * per-thread traversal stack in shared memory ([depth][thread]: conflict free) with
//     a local-memory overflow;
//   * nodes are 64 B fetched as 4 x ld.global.nc.v4, triangles 48 B as 3 x ld.global.nc.v4.
//
// Result semantics (DESIGN.md "traversal semantics", mirrored by oracle/o_scene.h):
// nearest Moller-Trumbore hit with t in (0, tmax), ties to the lower triangle id,
// back faces culled unless the instance is double sided, non-opaque candidates pass
// the hashed alpha test.  Box culling is conservative, so the result does not depend
// on tree shape, visiting order or scheduling.
#include "vkrt_device.cuh"
#include "vkrt_texture.cuh"
#include <cstdlib>
#include "vkrt_internal.h"

namespace vkrt {

#ifndef VKRT_STACK_SHARED
#define VKRT_STACK_SHARED 32
#endif
#ifndef VKRT_REFETCH
#define VKRT_REFETCH 24
#endif
#ifndef VKRT_TRACE_BLOCK
#define VKRT_TRACE_BLOCK 128
#endif
constexpr int      kTraceBlock       = VKRT_TRACE_BLOCK;
constexpr int      kStackShared      = VKRT_STACK_SHARED;  // entries per thread in shared memory
constexpr int      kStackLocal       = 128 - VKRT_STACK_SHARED;  // overflow entries in local memory
constexpr int      kRefetchThreshold = VKRT_REFETCH;  // refill the warp when fewer lanes than this are traversing
constexpr uint32_t kDone             = 0x7fffffffu;  // traversal finished (never a node index)
constexpr unsigned kFull             = 0xffffffffu;

enum TraceMode { MODE_CLOSEST = 0, MODE_SHADOW_RESOLVE = 1, MODE_ANY_PROBE = 2 };

// ---- alpha test for non-FORCE_OPAQUE instances (cold path) --------------------------------
__device__ __noinline__ bool alphaTestPasses(const SceneView& sc, uint32_t triId, float bu, float bv, uint32_t raySeed)
{
  const uint2              info = __ldg(&sc.triInfo[triId]);
  const InstanceXform&     I    = sc.instances[info.x];
  const InstanceData       geo  = sc.geoInfo[I.primMesh];
  const GltfShadeMaterial& mat  = sc.materials[max(0, geo.materialIndex)];
  float                    baseColorAlpha = mat.pbrBaseColorFactor[3];
  const int                texId = mat.pbrBaseColorTexture;
  if(texId > -1)
  {
    const uint32_t*         idx = reinterpret_cast<const uint32_t*>(geo.indexAddress) + 3ull * info.y;
    const VertexAttributes* vtx = reinterpret_cast<const VertexAttributes*>(geo.vertexAddress);
    const float*            t0  = vtx[idx[0]].texcoord;
    const float*            t1  = vtx[idx[1]].texcoord;
    const float*            t2  = vtx[idx[2]].texcoord;
    const float             b0  = 1.0f - bu - bv;
    float2 tc = f2(t0[0], t0[1]) * b0 + f2(t1[0], t1[1]) * bu + f2(t2[0], t2[1]) * bv;
    tc        = applyUvTransform(mat.uvTransform, tc);
    baseColorAlpha *= sampleTexture(sc.textures[texId], tc).w;
  }
  float opacity = (mat.alphaMode == VKRT_ALPHA_MASK) ? (baseColorAlpha > mat.alphaCutoff ? 1.0f : 0.0f) : baseColorAlpha;
  return !(alphaDraw(raySeed, triId) > opacity);
}

// One 64-byte node as two 256-bit loads (sm_100 LDG.E.256): the traversal is bound by L1 wavefronts of per-lane
// gathers, and two 32-byte requests per node cost half the tag/wavefront work of four 16-byte ones.
VKRT_D void loadNode256(const float4* __restrict__ nodes, uint32_t node, float4& n0, float4& n1, float4& n2, float4& n3)
{
  const float4* p = nodes + 4ull * node;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w)
      : "l"(p));
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w)
      : "l"(p + 2));
}

VKRT_D float safeInv(float x) { return 1.0f / (fabsf(x) < 1e-20f ? copysignf(1e-20f, x) : x); }

template <int MODE, bool COUNT, int VARIANT>
__global__ void __launch_bounds__(kTraceBlock, 8) traceKernel(const __grid_constant__ SceneView sc, const __grid_constant__ TraceLaunch L)
{
  __shared__ uint32_t s_stack[kStackShared * kTraceBlock];
  uint32_t* const sstack = &s_stack[threadIdx.x];
  uint32_t        lstack[kStackLocal];
  const uint32_t  n    = *L.count;
  const uint32_t  lane = threadIdx.x & 31u;
  const unsigned  lt   = (1u << lane) - 1u;

  if(sc.numTris == 0)
  {
    // nothing to hit: every closest ray misses, every shadow ray is unoccluded
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      const uint32_t slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? L.queue[i] : i;
      if(MODE == MODE_SHADOW_RESOLVE)
      {
        const float4 c = L.shC[i];
        float4       r = L.rad[__float_as_uint(c.w)];
        r.x += c.x; r.y += c.y; r.z += c.z;
        L.rad[__float_as_uint(c.w)] = r;
      }
      else
        L.hit[slot] = f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u));
    }
    return;
  }

  const float4* __restrict__ nodes = sc.nodes;
  const float4* __restrict__ tris  = sc.tris;

  // per-lane ray + traversal state (lives across refills)
  bool     haveRay = false;
  uint32_t slot = 0, seed = 0;
  float3   o = f3(0.f), d = f3(0.f), id = f3(0.f);
  float    tmax = 0.f, bestT = 0.f, bestU = 0.f, bestV = 0.f;
  uint32_t bestTri = ~0u;
  uint32_t node = kDone, leaf = 0;
  int      sp = 0;
  bool     exhausted = false;  // warp uniform: the queue has been handed out completely
  uint32_t nNodes = 0, nTris = 0;

  auto push = [&](uint32_t v) {
    if(sp < kStackShared)
      sstack[sp * kTraceBlock] = v;
    else
      lstack[min(sp - kStackShared, kStackLocal - 1)] = v;
    sp++;
  };
  auto pop = [&]() -> uint32_t {
    if(sp == 0)
      return kDone;
    --sp;
    return sp < kStackShared ? sstack[sp * kTraceBlock] : lstack[min(sp - kStackShared, kStackLocal - 1)];
  };

  for(;;)
  {
    __syncwarp();
    // ---- refill idle lanes ---------------------------------------------------------------
    const unsigned idle = __ballot_sync(kFull, !haveRay);
    if(idle && !exhausted)
    {
      const uint32_t want = (uint32_t)__popc(idle);
      uint32_t       base = 0;
      if(lane == 0)
        base = atomicAdd(L.cursor, want);
      base = __shfl_sync(kFull, base, 0);
      if(!haveRay)
      {
        const uint32_t i = base + (uint32_t)__popc(idle & lt);
        if(i < n)
        {
          slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? L.queue[i] : i;
          const float4 ro = L.rayO[slot], rd = L.rayD[slot];
          o = f3(ro); d = f3(rd);
          tmax = ro.w; seed = __float_as_uint(rd.w);
          id      = f3(safeInv(d.x), safeInv(d.y), safeInv(d.z));
          bestT   = tmax; bestU = 0.f; bestV = 0.f; bestTri = ~0u;
          node    = 0; leaf = 0; sp = 0;
          haveRay = true;
        }
      }
      exhausted = (base + want >= n);
    }
    if(!__any_sync(kFull, haveRay))
      break;

    if(haveRay)
    {
      while(node != kDone || leaf != 0)
      {
        // ---- inner nodes, until every lane of the warp has parked a leaf ---------------------
        while(node < kDone)  // inner index (leaf codes have bit 31 set, kDone is larger than any index)
        {
          float4 n0, n1, n2, n3;
          loadNode256(nodes, node, n0, n1, n2, n3);
          if(COUNT) nNodes++;
          const float a0x = (n0.x - o.x) * id.x, b0x = (n0.y - o.x) * id.x;
          const float a0y = (n0.z - o.y) * id.y, b0y = (n0.w - o.y) * id.y;
          const float a0z = (n2.x - o.z) * id.z, b0z = (n2.y - o.z) * id.z;
          const float a1x = (n1.x - o.x) * id.x, b1x = (n1.y - o.x) * id.x;
          const float a1y = (n1.z - o.y) * id.y, b1y = (n1.w - o.y) * id.y;
          const float a1z = (n2.z - o.z) * id.z, b1z = (n2.w - o.z) * id.z;
          const float tn0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
          const float tf0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), bestT));
          const float tn1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
          const float tf1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), bestT));
          const bool  h0  = tn0 <= tf0 * 1.0002f;  // conservative: never reject a box the triangle test could hit
          const bool  h1  = tn1 <= tf1 * 1.0002f;
          uint32_t    c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
          if(h0 && h1)
          {
            if(tn1 < tn0)
            {
              const uint32_t t = c0; c0 = c1; c1 = t;
            }
            push(c1);
            node = c0;
          }
          else if(h0)
            node = c0;
          else if(h1)
            node = c1;
          else
            node = pop();
          // park the first leaf and keep descending
          if((node & kLeafBit) && leaf == 0)
          {
            leaf = node;
            node = pop();
          }
          if(VARIANT == 0)
          {
            if(!__any_sync(__activemask(), leaf == 0))
              break;
          }
          else if(VARIANT == 1)
          {
            if(leaf != 0)
              break;  // if-if: handle the leaf right away
          }
          else
          {
            // leave as soon as only a few lanes are still searching
            if(__popc(__ballot_sync(__activemask(), leaf == 0)) <= VARIANT)
              break;
          }
        }

        // ---- parked leaves ---------------------------------------------------------------
        while(leaf != 0)
        {
          const uint32_t first = leaf & kLeafMask;
          const uint32_t cnt   = ((leaf >> kLeafCntSh) & 7u) + 1u;
          for(uint32_t k = 0; k < cnt; k++)
          {
            const float4 q0 = __ldg(&tris[3ull * (first + k) + 0]);
            const float4 q1 = __ldg(&tris[3ull * (first + k) + 1]);
            const float4 q2 = __ldg(&tris[3ull * (first + k) + 2]);
            if(COUNT) nTris++;
            const uint32_t flags = __float_as_uint(q0.w);
            const uint32_t triId = __float_as_uint(q1.w);
            // Moller-Trumbore, pinned operation order (no FMA: -fmad=false)
            const float3 v0  = f3(q0);
            const float3 e1  = f3(q1) - v0;
            const float3 e2  = f3(q2) - v0;
            const float3 p   = cross(d, e2);
            const float  det = dot(e1, p);
            if(!(flags & TRI_CULL_DISABLE))
            {
              const bool front = (flags & TRI_FLIP_FACING) ? (det < 0.0f) : (det > 0.0f);
              if(!front)
                continue;
            }
            if(det == 0.0f || !(det == det))
              continue;
            const float  inv = 1.0f / det;
            const float3 tv  = o - v0;
            const float  u   = dot(tv, p) * inv;
            if(!(u >= 0.0f && u <= 1.0f))
              continue;
            const float3 q = cross(tv, e1);
            const float  v = dot(d, q) * inv;
            if(!(v >= 0.0f && (u + v) <= 1.0f))
              continue;
            const float t = dot(e2, q) * inv;
            if(!(t > 0.0f && t < tmax))
              continue;
            if(!(t < bestT || (t == bestT && triId < bestTri)))
              continue;
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, u, v, seed))
              continue;
            bestT = t; bestU = u; bestV = v; bestTri = triId;
            if(MODE != MODE_CLOSEST)
            {
              node = kDone;  // TerminateOnFirstHit
              sp   = 0;
              break;
            }
          }
          if(node & kLeafBit)  // node may itself be a leaf popped while one was parked
          {
            leaf = node;
            node = pop();
          }
          else
            leaf = 0;
          if(MODE != MODE_CLOSEST && bestTri != ~0u)
          {
            leaf = 0;
            node = kDone;
          }
        }

        // ---- dynamic fetch: leave to refill when the warp has thinned out ------------------------
        if(!exhausted && __popc(__activemask()) < kRefetchThreshold)
          break;
      }

      if(node == kDone && leaf == 0)
      {
        if(MODE == MODE_CLOSEST)
          L.hit[slot] = (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(bestT, bestU, bestV, __uint_as_float(bestTri));
        else if(MODE == MODE_ANY_PROBE)
          L.hit[slot] = (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(0.f, 0.f, 0.f, __uint_as_float(0u));
        else if(bestTri == ~0u)
        {
          // unoccluded: radiance += contribution (one shadow ray per path per bounce: no race)
          const float4   c    = L.shC[slot];
          const uint32_t path = __float_as_uint(c.w);
          float4         rad  = L.rad[path];
          rad.x += c.x; rad.y += c.y; rad.z += c.z;
          L.rad[path] = rad;
        }
        haveRay = false;
      }
    }
  }

  if(COUNT && L.stats)
  {
    atomicAdd(&L.stats->nodesVisited, (unsigned long long)nNodes);
    atomicAdd(&L.stats->trisTested, (unsigned long long)nTris);
    if(MODE == MODE_SHADOW_RESOLVE)
    {
      atomicAdd(&L.stats->shadowNodes, (unsigned long long)nNodes);
      atomicAdd(&L.stats->shadowTris, (unsigned long long)nTris);
    }
  }
}


template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, 8) traceKernelIfIf(const __grid_constant__ SceneView sc, const __grid_constant__ TraceLaunch L)
{
  __shared__ uint32_t s_stack[kStackShared * kTraceBlock];
  uint32_t* const sstack = &s_stack[threadIdx.x];
  uint32_t        lstack[kStackLocal];
  const uint32_t  n    = *L.count;
  const uint32_t  lane = threadIdx.x & 31u;
  const unsigned  lt   = (1u << lane) - 1u;

  if(sc.numTris == 0)
  {
    // nothing to hit: every closest ray misses, every shadow ray is unoccluded
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      const uint32_t slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? L.queue[i] : i;
      if(MODE == MODE_SHADOW_RESOLVE)
      {
        const float4 c = L.shC[i];
        float4       r = L.rad[__float_as_uint(c.w)];
        r.x += c.x; r.y += c.y; r.z += c.z;
        L.rad[__float_as_uint(c.w)] = r;
      }
      else
        L.hit[slot] = f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u));
    }
    return;
  }

  const float4* __restrict__ nodes = sc.nodes;
  const float4* __restrict__ tris  = sc.tris;

  // per-lane ray + traversal state (lives across refills)
  bool     haveRay = false;
  uint32_t slot = 0, seed = 0;
  float3   o = f3(0.f), d = f3(0.f), id = f3(0.f);
  float    tmax = 0.f, bestT = 0.f, bestU = 0.f, bestV = 0.f;
  uint32_t bestTri = ~0u;
  uint32_t node = kDone, leaf = 0;
  int      sp = 0;
  bool     exhausted = false;  // warp uniform: the queue has been handed out completely
  uint32_t nNodes = 0, nTris = 0;

  auto push = [&](uint32_t v) {
    if(sp < kStackShared)
      sstack[sp * kTraceBlock] = v;
    else
      lstack[min(sp - kStackShared, kStackLocal - 1)] = v;
    sp++;
  };
  auto pop = [&]() -> uint32_t {
    if(sp == 0)
      return kDone;
    --sp;
    return sp < kStackShared ? sstack[sp * kTraceBlock] : lstack[min(sp - kStackShared, kStackLocal - 1)];
  };

  for(;;)
  {
    __syncwarp();
    // ---- refill idle lanes ---------------------------------------------------------------
    const unsigned idle = __ballot_sync(kFull, !haveRay);
    if(idle && !exhausted)
    {
      const uint32_t want = (uint32_t)__popc(idle);
      uint32_t       base = 0;
      if(lane == 0)
        base = atomicAdd(L.cursor, want);
      base = __shfl_sync(kFull, base, 0);
      if(!haveRay)
      {
        const uint32_t i = base + (uint32_t)__popc(idle & lt);
        if(i < n)
        {
          slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? L.queue[i] : i;
          const float4 ro = L.rayO[slot], rd = L.rayD[slot];
          o = f3(ro); d = f3(rd);
          tmax = ro.w; seed = __float_as_uint(rd.w);
          id      = f3(safeInv(d.x), safeInv(d.y), safeInv(d.z));
          bestT   = tmax; bestU = 0.f; bestV = 0.f; bestTri = ~0u;
          node    = 0; leaf = 0; sp = 0;
          haveRay = true;
        }
      }
      exhausted = (base + want >= n);
    }
    if(!__any_sync(kFull, haveRay))
      break;

    if(haveRay)
    {
      // single "if-if" loop: every iteration each lane does either one node step or one leaf
      while(node != kDone)
      {
        if(node < kDone)
        {
          float4 n0, n1, n2, n3;
          loadNode256(nodes, node, n0, n1, n2, n3);
          if(COUNT) nNodes++;
          const float a0x = (n0.x - o.x) * id.x, b0x = (n0.y - o.x) * id.x;
          const float a0y = (n0.z - o.y) * id.y, b0y = (n0.w - o.y) * id.y;
          const float a0z = (n2.x - o.z) * id.z, b0z = (n2.y - o.z) * id.z;
          const float a1x = (n1.x - o.x) * id.x, b1x = (n1.y - o.x) * id.x;
          const float a1y = (n1.z - o.y) * id.y, b1y = (n1.w - o.y) * id.y;
          const float a1z = (n2.z - o.z) * id.z, b1z = (n2.w - o.z) * id.z;
          const float tn0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
          const float tf0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), bestT));
          const float tn1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
          const float tf1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), bestT));
          const bool  h0  = tn0 <= tf0 * 1.0002f;
          const bool  h1  = tn1 <= tf1 * 1.0002f;
          uint32_t    c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
          if(h0 && h1)
          {
            if(tn1 < tn0)
            {
              const uint32_t t = c0; c0 = c1; c1 = t;
            }
            push(c1);
            node = c0;
          }
          else if(h0)
            node = c0;
          else if(h1)
            node = c1;
          else
            node = pop();
        }
        else
        {
          const uint32_t first = node & kLeafMask;
          const uint32_t cnt   = ((node >> kLeafCntSh) & 7u) + 1u;
          node = pop();
          for(uint32_t k = 0; k < cnt; k++)
          {
            const float4 q0 = __ldg(&tris[3ull * (first + k) + 0]);
            const float4 q1 = __ldg(&tris[3ull * (first + k) + 1]);
            const float4 q2 = __ldg(&tris[3ull * (first + k) + 2]);
            if(COUNT) nTris++;
            const uint32_t flags = __float_as_uint(q0.w);
            const uint32_t triId = __float_as_uint(q1.w);
            const float3 v0  = f3(q0);
            const float3 e1  = f3(q1) - v0;
            const float3 e2  = f3(q2) - v0;
            const float3 p   = cross(d, e2);
            const float  det = dot(e1, p);
            if(!(flags & TRI_CULL_DISABLE))
            {
              const bool front = (flags & TRI_FLIP_FACING) ? (det < 0.0f) : (det > 0.0f);
              if(!front)
                continue;
            }
            if(det == 0.0f || !(det == det))
              continue;
            const float  inv = 1.0f / det;
            const float3 tv  = o - v0;
            const float  u   = dot(tv, p) * inv;
            if(!(u >= 0.0f && u <= 1.0f))
              continue;
            const float3 q = cross(tv, e1);
            const float  v = dot(d, q) * inv;
            if(!(v >= 0.0f && (u + v) <= 1.0f))
              continue;
            const float t = dot(e2, q) * inv;
            if(!(t > 0.0f && t < tmax))
              continue;
            if(!(t < bestT || (t == bestT && triId < bestTri)))
              continue;
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, u, v, seed))
              continue;
            bestT = t; bestU = u; bestV = v; bestTri = triId;
            if(MODE != MODE_CLOSEST)
            {
              node = kDone;  // TerminateOnFirstHit
              sp   = 0;
              break;
            }
          }
        }
        if(!exhausted && __popc(__activemask()) < kRefetchThreshold)
          break;
      }

      if(node == kDone && leaf == 0)
      {
        if(MODE == MODE_CLOSEST)
          L.hit[slot] = (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(bestT, bestU, bestV, __uint_as_float(bestTri));
        else if(MODE == MODE_ANY_PROBE)
          L.hit[slot] = (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(0.f, 0.f, 0.f, __uint_as_float(0u));
        else if(bestTri == ~0u)
        {
          // unoccluded: radiance += contribution (one shadow ray per path per bounce: no race)
          const float4   c    = L.shC[slot];
          const uint32_t path = __float_as_uint(c.w);
          float4         rad  = L.rad[path];
          rad.x += c.x; rad.y += c.y; rad.z += c.z;
          L.rad[path] = rad;
        }
        haveRay = false;
      }
    }
  }

  if(COUNT && L.stats)
  {
    atomicAdd(&L.stats->nodesVisited, (unsigned long long)nNodes);
    atomicAdd(&L.stats->trisTested, (unsigned long long)nTris);
    if(MODE == MODE_SHADOW_RESOLVE)
    {
      atomicAdd(&L.stats->shadowNodes, (unsigned long long)nNodes);
      atomicAdd(&L.stats->shadowTris, (unsigned long long)nTris);
    }
  }
}



// ==========================================================================================================
// Warp-cooperative variant ("leaf queue").  Lanes own rays for the box tests, but triangle tests are handed to
// a warp-shared queue in shared memory and executed 32 at a time by whichever lanes are free, so neither phase
// runs with a handful of active lanes:
//   node phase   every lane whose ray still has inner nodes does one node step; leaf children are appended to
//                the queue as (owner lane, leaf code); the lane keeps descending with its current best t
//                (speculative: best t only shrinks, so stale values just cost extra visits);
//   leaf phase   when >= 32 items wait (or the node phase runs dry) each lane takes one item, reads the owner's
//                ray through warp shuffles and resolves the result with a 64-bit shared-memory atomicMin on
//                (t bits << 32 | triangle id) -- exactly the "nearest t, ties to the lower id" rule.
// A ray is finished when its stack is empty and none of its items are pending; finished lanes are refilled from
// the global cursor.  Results are identical to the per-lane kernel (acceptance is order independent).
// ==========================================================================================================
constexpr int kQCap     = 128;  // leaf-queue capacity per warp
constexpr int kNodeIter = 12;   // node steps between refill checks

struct WarpShared
{
  unsigned long long bestKey[32];
  float2             bestUV[32];
  int                pend[32];
  uint2              queue[kQCap];
};

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, 8) traceKernelCoop(const __grid_constant__ SceneView sc, const __grid_constant__ TraceLaunch L)
{
  __shared__ uint32_t   s_stack[kStackShared * kTraceBlock];
  __shared__ WarpShared s_ws[kTraceBlock / 32];
  uint32_t* const sstack = &s_stack[threadIdx.x];
  WarpShared&     ws     = s_ws[threadIdx.x >> 5];
  uint32_t        lstack[kStackLocal];
  const uint32_t  n    = *L.count;
  const uint32_t  lane = threadIdx.x & 31u;
  const unsigned  lt   = (1u << lane) - 1u;

  if(sc.numTris == 0)
  {
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      const uint32_t slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? L.queue[i] : i;
      if(MODE == MODE_SHADOW_RESOLVE)
      {
        const float4 c = L.shC[i];
        float4       r = L.rad[__float_as_uint(c.w)];
        r.x += c.x; r.y += c.y; r.z += c.z;
        L.rad[__float_as_uint(c.w)] = r;
      }
      else
        L.hit[slot] = f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u));
    }
    return;
  }

  const float4* __restrict__ nodes = sc.nodes;
  const float4* __restrict__ tris  = sc.tris;

  bool     haveRay = false;
  uint32_t slot = 0, seed = 0;
  float3   o = f3(0.f), d = f3(0.f), id = f3(0.f);
  float    tmax = 0.f;
  uint32_t node = kDone;
  int      sp = 0;
  int      qcount = 0;          // warp uniform
  bool     exhausted = false;   // warp uniform
  uint32_t nNodes = 0, nTris = 0;
  ws.pend[lane] = 0;

  auto push = [&](uint32_t v) {
    if(sp < kStackShared)
      sstack[sp * kTraceBlock] = v;
    else
      lstack[min(sp - kStackShared, kStackLocal - 1)] = v;
    sp++;
  };
  auto pop = [&]() -> uint32_t {
    if(sp == 0)
      return kDone;
    --sp;
    return sp < kStackShared ? sstack[sp * kTraceBlock] : lstack[min(sp - kStackShared, kStackLocal - 1)];
  };

  for(;;)
  {
    __syncwarp();
    // ---- refill idle lanes -----------------------------------------------------------------------
    const unsigned idle = __ballot_sync(kFull, !haveRay);
    if(idle && !exhausted)
    {
      const uint32_t want = (uint32_t)__popc(idle);
      uint32_t       base = 0;
      if(lane == 0)
        base = atomicAdd(L.cursor, want);
      base = __shfl_sync(kFull, base, 0);
      if(!haveRay)
      {
        const uint32_t i = base + (uint32_t)__popc(idle & lt);
        if(i < n)
        {
          slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? L.queue[i] : i;
          const float4 ro = L.rayO[slot], rd = L.rayD[slot];
          o = f3(ro); d = f3(rd);
          tmax = ro.w; seed = __float_as_uint(rd.w);
          id   = f3(safeInv(d.x), safeInv(d.y), safeInv(d.z));
          ws.bestKey[lane] = ((unsigned long long)__float_as_uint(tmax) << 32) | 0xffffffffull;
          ws.pend[lane]    = 0;
          node = 0; sp = 0;
          haveRay = true;
        }
      }
      exhausted = (base + want >= n);
    }
    if(!__any_sync(kFull, haveRay))
      break;
    __syncwarp();

    // ---- node phase ---------------------------------------------------------------------------------
    for(int it = 0; it < kNodeIter; it++)
    {
      const unsigned long long myKey = ws.bestKey[lane];
      const bool found = (MODE != MODE_CLOSEST) && haveRay && (uint32_t)myKey != 0xffffffffu;  // any-hit: stop descending
      const bool work  = haveRay && node != kDone && !found;
      const unsigned wmask = __ballot_sync(kFull, work);
      if(wmask == 0 || qcount > kQCap - 64)
        break;
      bool     l0 = false, l1 = false;
      uint32_t c0 = 0, c1 = 0;
      if(work)
      {
        const float  bestT = __uint_as_float((uint32_t)(myKey >> 32));
        float4 n0, n1, n2, n3;
        loadNode256(nodes, node, n0, n1, n2, n3);
        if(COUNT) nNodes++;
        const float a0x = (n0.x - o.x) * id.x, b0x = (n0.y - o.x) * id.x;
        const float a0y = (n0.z - o.y) * id.y, b0y = (n0.w - o.y) * id.y;
        const float a0z = (n2.x - o.z) * id.z, b0z = (n2.y - o.z) * id.z;
        const float a1x = (n1.x - o.x) * id.x, b1x = (n1.y - o.x) * id.x;
        const float a1y = (n1.z - o.y) * id.y, b1y = (n1.w - o.y) * id.y;
        const float a1z = (n2.z - o.z) * id.z, b1z = (n2.w - o.z) * id.z;
        const float tn0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
        const float tf0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), bestT));
        const float tn1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
        const float tf1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), bestT));
        const bool  h0  = tn0 <= tf0 * 1.0002f;  // conservative: never reject a box the triangle test could hit
        const bool  h1  = tn1 <= tf1 * 1.0002f;
        c0 = __float_as_uint(n3.x);
        c1 = __float_as_uint(n3.y);
        l0 = h0 && (c0 & kLeafBit);
        l1 = h1 && (c1 & kLeafBit);
        const bool i0 = h0 && !(c0 & kLeafBit);
        const bool i1 = h1 && !(c1 & kLeafBit);
        if(i0 && i1)
        {
          if(tn1 < tn0)
          {
            push(c0);
            node = c1;
          }
          else
          {
            push(c1);
            node = c0;
          }
        }
        else if(i0)
          node = c0;
        else if(i1)
          node = c1;
        else
          node = pop();
      }
      // append leaf children to the warp queue (uniform code)
      const unsigned m0 = __ballot_sync(kFull, l0);
      const unsigned m1 = __ballot_sync(kFull, l1);
      if(l0)
        ws.queue[qcount + __popc(m0 & lt)] = make_uint2(lane, c0);
      if(l1)
        ws.queue[qcount + __popc(m0) + __popc(m1 & lt)] = make_uint2(lane, c1);
      if(l0 || l1)
        ws.pend[lane] += (l0 ? 1 : 0) + (l1 ? 1 : 0);
      qcount += __popc(m0) + __popc(m1);
      if(qcount >= 32 && __popc(wmask) < 24)
        break;  // enough triangle work for a full batch and the node phase is thinning out
    }
    __syncwarp();

    // ---- leaf phase -----------------------------------------------------------------------------------
    {
      const unsigned stillWork = __ballot_sync(kFull, haveRay && node != kDone);
      // drain in full batches; take the remainder too when the node phase has (almost) nothing left
      while(qcount >= 32 || (qcount > 0 && (__popc(stillWork) < 16 || qcount > kQCap - 64)))
      {
        const int  batch = min(32, qcount);
        const bool valid = (int)lane < batch;
        uint2      item  = make_uint2(lane, 0u);
        if(valid)
          item = ws.queue[qcount - batch + lane];
        const int   owner = (int)item.x;
        const float rox = __shfl_sync(kFull, o.x, owner), roy = __shfl_sync(kFull, o.y, owner), roz = __shfl_sync(kFull, o.z, owner);
        const float rdx = __shfl_sync(kFull, d.x, owner), rdy = __shfl_sync(kFull, d.y, owner), rdz = __shfl_sync(kFull, d.z, owner);
        const float rtmax = __shfl_sync(kFull, tmax, owner);
        const uint32_t rseed = __shfl_sync(kFull, seed, owner);
        unsigned long long cand = ~0ull;
        float cu = 0.f, cv = 0.f;
        if(valid)
        {
          const float3 ro = f3(rox, roy, roz), rd = f3(rdx, rdy, rdz);
          const uint32_t first = item.y & kLeafMask;
          const uint32_t cnt   = ((item.y >> kLeafCntSh) & 7u) + 1u;
          unsigned long long cur = ws.bestKey[owner];   // may be stale: only used to skip work
          for(uint32_t k = 0; k < cnt; k++)
          {
            const float4 q0 = __ldg(&tris[3ull * (first + k) + 0]);
            const float4 q1 = __ldg(&tris[3ull * (first + k) + 1]);
            const float4 q2 = __ldg(&tris[3ull * (first + k) + 2]);
            if(COUNT) nTris++;
            const uint32_t flags = __float_as_uint(q0.w);
            const uint32_t triId = __float_as_uint(q1.w);
            // Moller-Trumbore, pinned operation order (no FMA: -fmad=false)
            const float3 v0  = f3(q0);
            const float3 e1  = f3(q1) - v0;
            const float3 e2  = f3(q2) - v0;
            const float3 p   = cross(rd, e2);
            const float  det = dot(e1, p);
            if(!(flags & TRI_CULL_DISABLE))
            {
              const bool front = (flags & TRI_FLIP_FACING) ? (det < 0.0f) : (det > 0.0f);
              if(!front)
                continue;
            }
            if(det == 0.0f || !(det == det))
              continue;
            const float  inv = 1.0f / det;
            const float3 tv  = ro - v0;
            const float  u   = dot(tv, p) * inv;
            if(!(u >= 0.0f && u <= 1.0f))
              continue;
            const float3 q = cross(tv, e1);
            const float  v = dot(rd, q) * inv;
            if(!(v >= 0.0f && (u + v) <= 1.0f))
              continue;
            const float t = dot(e2, q) * inv;
            if(!(t > 0.0f && t < rtmax))
              continue;
            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | triId;
            if(!(key < cur && key < cand))
              continue;
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, u, v, rseed))
              continue;
            cand = key; cu = u; cv = v;
          }
          if(cand != ~0ull)
            atomicMin(&ws.bestKey[owner], cand);
        }
        __syncwarp();
        if(valid)
        {
          if(cand != ~0ull && ws.bestKey[owner] == cand)
            ws.bestUV[owner] = make_float2(cu, cv);
          atomicSub(&ws.pend[owner], 1);
        }
        __syncwarp();
        qcount -= batch;
      }
    }

    // ---- finished rays -----------------------------------------------------------------------------------
    if(haveRay)
    {
      const unsigned long long key = ws.bestKey[lane];
      const bool hitFound = (uint32_t)key != 0xffffffffu;
      const bool finished = ((node == kDone) || (MODE != MODE_CLOSEST && hitFound)) && ws.pend[lane] == 0;
      if(finished)
      {
        if(MODE == MODE_CLOSEST)
        {
          const float2 uv = ws.bestUV[lane];
          L.hit[slot] = hitFound ? f4(__uint_as_float((uint32_t)(key >> 32)), uv.x, uv.y, __uint_as_float((uint32_t)key))
                                 : f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u));
        }
        else if(MODE == MODE_ANY_PROBE)
          L.hit[slot] = hitFound ? f4(0.f, 0.f, 0.f, __uint_as_float(0u)) : f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u));
        else if(!hitFound)
        {
          const float4   c    = L.shC[slot];
          const uint32_t path = __float_as_uint(c.w);
          float4         rad  = L.rad[path];
          rad.x += c.x; rad.y += c.y; rad.z += c.z;
          L.rad[path] = rad;
        }
        haveRay = false;
        node    = kDone;
        sp      = 0;
      }
    }
  }

  if(COUNT && L.stats)
  {
    atomicAdd(&L.stats->nodesVisited, (unsigned long long)nNodes);
    atomicAdd(&L.stats->trisTested, (unsigned long long)nTris);
    if(MODE == MODE_SHADOW_RESOLVE)
    {
      atomicAdd(&L.stats->shadowNodes, (unsigned long long)nNodes);
      atomicAdd(&L.stats->shadowTris, (unsigned long long)nTris);
    }
  }
}

template <int MODE, bool COUNT>
static void launchTraceCoop(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  static int grid = 0;
  if(!grid)
  {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, traceKernelCoop<MODE, COUNT>, kTraceBlock, 0);
    grid = numSMs * (perSM < 1 ? 1 : perSM);
  }
  traceKernelCoop<MODE, COUNT><<<grid, kTraceBlock, 0, s>>>(scene, L);
}

template <int MODE, bool COUNT>
static void launchTraceIfIf(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  static int grid = 0;
  if(!grid)
  {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, traceKernelIfIf<MODE, COUNT>, kTraceBlock, 0);
    grid = numSMs * (perSM < 1 ? 1 : perSM);
  }
  traceKernelIfIf<MODE, COUNT><<<grid, kTraceBlock, 0, s>>>(scene, L);
}

template <int MODE, bool COUNT, int VARIANT>
static void launchTraceV(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  static int grid = 0;
  if(!grid)
  {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, traceKernel<MODE, COUNT, VARIANT>, kTraceBlock, 0);
    grid = numSMs * (perSM < 1 ? 1 : perSM);
  }
  traceKernel<MODE, COUNT, VARIANT><<<grid, kTraceBlock, 0, s>>>(scene, L);
}
template <int MODE, bool COUNT>
static void launchTrace(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  static int variant = -1;
  if(variant < 0)
  {
    const char* e = getenv("VKRT_TRACE_VARIANT");
    variant       = e ? atoi(e) : 8;  // default: leave the node loop when <= 8 lanes are still searching
  }
  switch(variant)
  {
    case 100: launchTraceCoop<MODE, COUNT>(s, scene, L, numSMs); break;
    case 200: launchTraceIfIf<MODE, COUNT>(s, scene, L, numSMs); break;
    case 1: launchTraceV<MODE, COUNT, 1>(s, scene, L, numSMs); break;
    case 4: launchTraceV<MODE, COUNT, 4>(s, scene, L, numSMs); break;
    case 8: launchTraceV<MODE, COUNT, 8>(s, scene, L, numSMs); break;
    case 16: launchTraceV<MODE, COUNT, 16>(s, scene, L, numSMs); break;
    default: launchTraceV<MODE, COUNT, 0>(s, scene, L, numSMs); break;
  }
}

void launchTraceClosest(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, bool count, int numSMs)
{
  if(count)
    launchTrace<MODE_CLOSEST, true>(s, scene, L, numSMs);
  else
    launchTrace<MODE_CLOSEST, false>(s, scene, L, numSMs);
}
void launchTraceShadow(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, bool count, int numSMs)
{
  if(count)
    launchTrace<MODE_SHADOW_RESOLVE, true>(s, scene, L, numSMs);
  else
    launchTrace<MODE_SHADOW_RESOLVE, false>(s, scene, L, numSMs);
}
void launchTraceAnyProbe(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  launchTrace<MODE_ANY_PROBE, false>(s, scene, L, numSMs);
}

}  // namespace vkrt
