// vkrt_trace.cu -- persistent-thread BVH2 traversal kernels for sm_100a.
//
// Replaces the driver traversal behind ClosestHit()/AnyHit() of the reference
// (shaders/traceray_rq.glsl:108-185, shaders/traceray_rtx.glsl:30-73) and the inline
// alpha test HitTest()/pathtrace.rahit (traceray_rq.glsl:32-102).
//
// Kernel shape: grid = SMs x resident CTAs (persistent), each warp pulls batches of
// 32 ray slots from a global cursor; per-thread traversal stack lives in shared
// memory ([depth][thread] so a warp's accesses are conflict free) with a local-memory
// overflow; nodes are 64 B fetched as 4 x ld.global.nc.v4 (read-only path), triangles
// 48 B as 3 x ld.global.nc.v4.
//
// Result semantics (DESIGN.md "traversal semantics", mirrored by oracle/o_scene.h):
// nearest Moller-Trumbore hit with t in (0, tmax), ties to the lower triangle id,
// back faces culled unless the instance is double sided, non-opaque candidates pass
// the hashed alpha test.  Box culling is conservative, so the result does not depend
// on tree shape or visiting order.
#include "vkrt_device.cuh"
#include "vkrt_texture.cuh"
#include "vkrt_internal.h"

namespace vkrt {

constexpr int kTraceBlock   = 128;
constexpr int kStackShared  = 32;   // entries per thread in shared memory
constexpr int kStackLocal   = 96;   // overflow entries in local memory

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

VKRT_D float safeInv(float x) { return 1.0f / (fabsf(x) < 1e-20f ? copysignf(1e-20f, x) : x); }

struct TravResult
{
  float    t, u, v;
  uint32_t tri;
};

// One ray through the BVH.  `sstack` points at this thread's column of the shared stack.
template <bool CLOSEST, bool COUNT>
VKRT_D TravResult traverse(const SceneView& sc, float3 o, float3 d, float tmax, uint32_t seed, uint32_t* sstack, uint32_t& nNodes, uint32_t& nTris)
{
  TravResult best{tmax, 0.f, 0.f, ~0u};
  if(sc.numTris == 0)
    return best;
  const float idx = safeInv(d.x), idy = safeInv(d.y), idz = safeInv(d.z);
  uint32_t    lstack[kStackLocal];
  int         sp   = 0;
  uint32_t    node = 0;
  const float4* __restrict__ nodes = sc.nodes;
  const float4* __restrict__ tris  = sc.tris;

  for(;;)
  {
    // ---- inner node: test both children ---------------------------------------------------
    const float4 n0 = __ldg(&nodes[4ull * node + 0]);
    const float4 n1 = __ldg(&nodes[4ull * node + 1]);
    const float4 n2 = __ldg(&nodes[4ull * node + 2]);
    const float4 n3 = __ldg(&nodes[4ull * node + 3]);
    if(COUNT) nNodes++;
    const float a0x = (n0.x - o.x) * idx, b0x = (n0.y - o.x) * idx;
    const float a0y = (n0.z - o.y) * idy, b0y = (n0.w - o.y) * idy;
    const float a0z = (n2.x - o.z) * idz, b0z = (n2.y - o.z) * idz;
    const float a1x = (n1.x - o.x) * idx, b1x = (n1.y - o.x) * idx;
    const float a1y = (n1.z - o.y) * idy, b1y = (n1.w - o.y) * idy;
    const float a1z = (n2.z - o.z) * idz, b1z = (n2.w - o.z) * idz;
    const float tn0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
    const float tf0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), best.t));
    const float tn1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
    const float tf1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), best.t));
    const bool  h0  = tn0 <= tf0 * 1.0002f;  // conservative: never reject a box the triangle test could hit
    const bool  h1  = tn1 <= tf1 * 1.0002f;
    uint32_t    c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
    uint32_t    cur;
    if(h0 && h1)
    {
      if(tn1 < tn0)
      {
        uint32_t tmp = c0; c0 = c1; c1 = tmp;
      }
      if(sp < kStackShared)
        sstack[sp * kTraceBlock] = c1;
      else
        lstack[min(sp - kStackShared, kStackLocal - 1)] = c1;
      sp++;
      cur = c0;
    }
    else if(h0)
      cur = c0;
    else if(h1)
      cur = c1;
    else
    {
      if(sp == 0)
        return best;
      --sp;
      cur = sp < kStackShared ? sstack[sp * kTraceBlock] : lstack[min(sp - kStackShared, kStackLocal - 1)];
    }

    // ---- leaves ---------------------------------------------------------------------------------
    while(cur & kLeafBit)
    {
      const uint32_t first = cur & kLeafMask;
      const uint32_t cnt   = ((cur >> kLeafCntSh) & 7u) + 1u;
      for(uint32_t k = 0; k < cnt; k++)
      {
        const float4 q0 = __ldg(&tris[3ull * (first + k) + 0]);
        const float4 q1 = __ldg(&tris[3ull * (first + k) + 1]);
        const float4 q2 = __ldg(&tris[3ull * (first + k) + 2]);
        if(COUNT) nTris++;
        const uint32_t flags = __float_as_uint(q0.w);
        const uint32_t triId = __float_as_uint(q1.w);
        // Moller-Trumbore, pinned operation order (no FMA: -fmad=false)
        const float3 v0 = f3(q0);
        const float3 e1 = f3(q1) - v0;
        const float3 e2 = f3(q2) - v0;
        const float3 p  = cross(d, e2);
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
        if(!(t < best.t || (t == best.t && triId < best.tri)))
          continue;
        if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, u, v, seed))
          continue;
        best.t = t; best.u = u; best.v = v; best.tri = triId;
        if(!CLOSEST)
          return best;  // TerminateOnFirstHit
      }
      if(sp == 0)
        return best;
      --sp;
      cur = sp < kStackShared ? sstack[sp * kTraceBlock] : lstack[min(sp - kStackShared, kStackLocal - 1)];
    }
    node = cur;
  }
}

// ---- closest hit -------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock) traceClosestKernel(const __grid_constant__ SceneView sc, const __grid_constant__ TraceLaunch L)
{
  __shared__ uint32_t s_stack[kStackShared * kTraceBlock];
  uint32_t*      sstack = &s_stack[threadIdx.x];
  const uint32_t n      = *L.count;
  const uint32_t lane   = threadIdx.x & 31u;
  uint32_t       nNodes = 0, nTris = 0, nRays = 0;
  for(;;)
  {
    uint32_t base = 0;
    if(lane == 0)
      base = atomicAdd(L.cursor, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if(base >= n)
      break;
    const uint32_t i = base + lane;
    if(i < n)
    {
      const uint32_t slot = L.queue ? L.queue[i] : i;
      const float4   ro = L.rayO[slot], rd = L.rayD[slot];
      TravResult r = traverse<true, COUNT>(sc, f3(ro), f3(rd), ro.w, __float_as_uint(rd.w), sstack, nNodes, nTris);
      L.hit[slot]  = (r.tri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(r.t, r.u, r.v, __uint_as_float(r.tri));
      nRays++;
    }
  }
  if(COUNT && L.stats)
  {
    atomicAdd(&L.stats->nodesVisited, (unsigned long long)nNodes);
    atomicAdd(&L.stats->trisTested, (unsigned long long)nTris);
  }
  (void)nRays;
}

// ---- shadow (any hit) + resolve -------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock) traceShadowKernel(const __grid_constant__ SceneView sc, const __grid_constant__ TraceLaunch L)
{
  __shared__ uint32_t s_stack[kStackShared * kTraceBlock];
  uint32_t*      sstack = &s_stack[threadIdx.x];
  const uint32_t n      = *L.count;
  const uint32_t lane   = threadIdx.x & 31u;
  uint32_t       nNodes = 0, nTris = 0;
  for(;;)
  {
    uint32_t base = 0;
    if(lane == 0)
      base = atomicAdd(L.cursor, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if(base >= n)
      break;
    const uint32_t i = base + lane;
    if(i < n)
    {
      const float4 ro = L.rayO[i], rd = L.rayD[i];
      TravResult   r  = traverse<false, COUNT>(sc, f3(ro), f3(rd), ro.w, __float_as_uint(rd.w), sstack, nNodes, nTris);
      if(r.tri == ~0u)
      {
        // unoccluded: radiance += contribution (one shadow ray per path per bounce: no race)
        const float4   c    = L.shC[i];
        const uint32_t path = __float_as_uint(c.w);
        float4         rad  = L.rad[path];
        rad.x += c.x; rad.y += c.y; rad.z += c.z;
        L.rad[path] = rad;
      }
    }
  }
  if(COUNT && L.stats)
  {
    atomicAdd(&L.stats->nodesVisited, (unsigned long long)nNodes);
    atomicAdd(&L.stats->trisTested, (unsigned long long)nTris);
    atomicAdd(&L.stats->shadowNodes, (unsigned long long)nNodes);
    atomicAdd(&L.stats->shadowTris, (unsigned long long)nTris);
  }
}

// ---- probe: explicit ray list (tests, picking) -------------------------------------------------------
__global__ void __launch_bounds__(kTraceBlock) traceProbeKernel(const __grid_constant__ SceneView sc, const float4* __restrict__ rays, uint32_t n,
                                                                int closest, float4* __restrict__ out)
{
  __shared__ uint32_t s_stack[kStackShared * kTraceBlock];
  uint32_t* sstack = &s_stack[threadIdx.x];
  uint32_t  nNodes = 0, nTris = 0;
  for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const float4 ro = rays[2ull * i], rd = rays[2ull * i + 1];
    if(closest)
    {
      TravResult r = traverse<true, false>(sc, f3(ro), f3(rd), ro.w, __float_as_uint(rd.w), sstack, nNodes, nTris);
      out[i]       = (r.tri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(r.t, r.u, r.v, __uint_as_float(r.tri));
    }
    else
    {
      TravResult r = traverse<false, false>(sc, f3(ro), f3(rd), ro.w, __float_as_uint(rd.w), sstack, nNodes, nTris);
      out[i]       = (r.tri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(0.f, 0.f, 0.f, __uint_as_float(0u));
    }
  }
}

static int traceGrid(int numSMs, const void* fn)
{
  int perSM = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, fn, kTraceBlock, 0);
  if(perSM < 1)
    perSM = 1;
  return numSMs * perSM;
}

void launchTraceClosest(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, bool count, int numSMs)
{
  static int gridPlain = 0, gridCount = 0;
  if(count)
  {
    if(!gridCount) gridCount = traceGrid(numSMs, (const void*)traceClosestKernel<true>);
    traceClosestKernel<true><<<gridCount, kTraceBlock, 0, s>>>(scene, L);
  }
  else
  {
    if(!gridPlain) gridPlain = traceGrid(numSMs, (const void*)traceClosestKernel<false>);
    traceClosestKernel<false><<<gridPlain, kTraceBlock, 0, s>>>(scene, L);
  }
}
void launchTraceShadow(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, bool count, int numSMs)
{
  static int gridPlain = 0, gridCount = 0;
  if(count)
  {
    if(!gridCount) gridCount = traceGrid(numSMs, (const void*)traceShadowKernel<true>);
    traceShadowKernel<true><<<gridCount, kTraceBlock, 0, s>>>(scene, L);
  }
  else
  {
    if(!gridPlain) gridPlain = traceGrid(numSMs, (const void*)traceShadowKernel<false>);
    traceShadowKernel<false><<<gridPlain, kTraceBlock, 0, s>>>(scene, L);
  }
}
void launchTraceProbe(cudaStream_t s, const SceneView& scene, const float4* rays, uint32_t n, bool closest, float4* out)
{
  uint32_t blocks = (n + kTraceBlock - 1) / kTraceBlock;
  int      grid   = (int)(blocks > 4096u ? 4096u : blocks);
  if(grid < 1) grid = 1;
  traceProbeKernel<<<grid, kTraceBlock, 0, s>>>(scene, rays, n, closest ? 1 : 0, out);
}

}  // namespace vkrt
