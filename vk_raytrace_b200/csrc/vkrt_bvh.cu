// vkrt_bvh.cu -- GPU acceleration-structure build for sm_100a.
//
// Replaces AccelStructure::createBottomLevelAS / createTopLevelAS of the reference
// (src/accelstruct.cpp:110-162), where the Vulkan driver builds an opaque two-level
// BVH.  B200 design: 180 GB of HBM makes instancing-by-reference unnecessary for the
// scene sizes of this path, so every instance is flattened to world space once and a
// single-level BVH2 is built over all triangles:
//   1. flatten      instance x triangle -> world-space triangle, bounds, centroid
//   2. morton       63-bit Morton code of the centroid
//   3. sort         radix sort (CUB) of (code, triangle id)
//   4. hierarchy    Karras 2012 binary radix tree, one thread per internal node
//   5. refit        bottom-up AABB + SAH cost with atomic arrival flags
//   6. emit         64-byte two-child-box nodes, optional SAH leaf collapse
// Instance flag semantics (FORCE_OPAQUE, CULL_DISABLE) are carried per triangle.
#include <cub/cub.cuh>
#include <cfloat>
#include "vkrt_device.cuh"
#include "vkrt_internal.h"

namespace vkrt {

__device__ __forceinline__ uint32_t orderedFromFloat(float f)
{
  uint32_t u = __float_as_uint(f);
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float floatFromOrdered(uint32_t u)
{
  u ^= ((u >> 31) - 1u) | 0x80000000u;
  return __uint_as_float(u);
}

__device__ __forceinline__ uint64_t expandBits21(uint32_t v)
{
  uint64_t x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

// ---- 1. flatten ----------------------------------------------------------------------
__global__ void flattenKernel(const InstanceXform* __restrict__ inst, const uint32_t* __restrict__ instTriOffset, uint32_t numInst,
                              const InstanceData* __restrict__ geo, uint32_t numTris, float4* __restrict__ wtri,
                              uint2* __restrict__ triInfo, uint4* __restrict__ triShade, uint32_t* __restrict__ boundsOrdered /*6*/)
{
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  float3   c   = f3(0.f);
  bool     ok  = gid < numTris;
  if(ok)
  {
    // instance lookup: last i with instTriOffset[i] <= gid
    uint32_t lo = 0, hi = numInst;
    while(hi - lo > 1)
    {
      uint32_t mid = (lo + hi) >> 1;
      if(instTriOffset[mid] <= gid)
        lo = mid;
      else
        hi = mid;
    }
    const InstanceXform& I    = inst[lo];
    uint32_t             prim = gid - I.triOffset;
    const InstanceData   g    = geo[I.primMesh];
    const uint32_t*      idx  = reinterpret_cast<const uint32_t*>(g.indexAddress) + 3ull * prim;
    const VertexAttributes* vtx = reinterpret_cast<const VertexAttributes*>(g.vertexAddress);
    const float* pa = vtx[idx[0]].position;
    const float* pb = vtx[idx[1]].position;
    const float* pc = vtx[idx[2]].position;
    float3 v0 = xformPoint(I.o2w, f3(pa[0], pa[1], pa[2]));
    float3 v1 = xformPoint(I.o2w, f3(pb[0], pb[1], pb[2]));
    float3 v2 = xformPoint(I.o2w, f3(pc[0], pc[1], pc[2]));
    wtri[3ull * gid + 0] = f4(v0, __uint_as_float(I.flags));
    wtri[3ull * gid + 1] = f4(v1, __uint_as_float(gid));
    wtri[3ull * gid + 2] = f4(v2, __uint_as_float((uint32_t)max(0, g.materialIndex)));  // for the any-hit alpha test
    triInfo[gid]         = make_uint2(lo, prim);
    // per-triangle shading record: the three 32-byte vertices + {instance, material}; one dependent load from the hit
    // instead of triInfo -> instance -> geometry -> indices -> vertices
    {
      const uint4* v4 = reinterpret_cast<const uint4*>(vtx);
      uint4*       ts = triShade + 7ull * gid;
      ts[0] = v4[2ull * idx[0]]; ts[1] = v4[2ull * idx[0] + 1];
      ts[2] = v4[2ull * idx[1]]; ts[3] = v4[2ull * idx[1] + 1];
      ts[4] = v4[2ull * idx[2]]; ts[5] = v4[2ull * idx[2] + 1];
      ts[6] = make_uint4(lo, (uint32_t)max(0, g.materialIndex), I.flags, prim);
    }
    float3 lo3 = f3(fminf(v0.x, fminf(v1.x, v2.x)), fminf(v0.y, fminf(v1.y, v2.y)), fminf(v0.z, fminf(v1.z, v2.z)));
    float3 hi3 = f3(fmaxf(v0.x, fmaxf(v1.x, v2.x)), fmaxf(v0.y, fmaxf(v1.y, v2.y)), fmaxf(v0.z, fmaxf(v1.z, v2.z)));
    c          = (lo3 + hi3) * 0.5f;
  }
  // centroid bounds: warp reduce then one atomic per warp
  float mnx = ok ? c.x : FLT_MAX, mny = ok ? c.y : FLT_MAX, mnz = ok ? c.z : FLT_MAX;
  float mxx = ok ? c.x : -FLT_MAX, mxy = ok ? c.y : -FLT_MAX, mxz = ok ? c.z : -FLT_MAX;
#pragma unroll
  for(int o = 16; o > 0; o >>= 1)
  {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mnz = fminf(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    mxz = fmaxf(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
  }
  if((threadIdx.x & 31) == 0 && mnx <= mxx)
  {
    atomicMin(&boundsOrdered[0], orderedFromFloat(mnx));
    atomicMin(&boundsOrdered[1], orderedFromFloat(mny));
    atomicMin(&boundsOrdered[2], orderedFromFloat(mnz));
    atomicMax(&boundsOrdered[3], orderedFromFloat(mxx));
    atomicMax(&boundsOrdered[4], orderedFromFloat(mxy));
    atomicMax(&boundsOrdered[5], orderedFromFloat(mxz));
  }
}

// ---- 2. morton ------------------------------------------------------------------------
__global__ void mortonKernel(const float4* __restrict__ wtri, uint32_t numTris, const uint32_t* __restrict__ boundsOrdered,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
  uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if(gid >= numTris)
    return;
  float3 bmin = f3(floatFromOrdered(boundsOrdered[0]), floatFromOrdered(boundsOrdered[1]), floatFromOrdered(boundsOrdered[2]));
  float3 bmax = f3(floatFromOrdered(boundsOrdered[3]), floatFromOrdered(boundsOrdered[4]), floatFromOrdered(boundsOrdered[5]));
  float3 v0 = f3(wtri[3ull * gid]), v1 = f3(wtri[3ull * gid + 1]), v2 = f3(wtri[3ull * gid + 2]);
  float3 lo3 = f3(fminf(v0.x, fminf(v1.x, v2.x)), fminf(v0.y, fminf(v1.y, v2.y)), fminf(v0.z, fminf(v1.z, v2.z)));
  float3 hi3 = f3(fmaxf(v0.x, fmaxf(v1.x, v2.x)), fmaxf(v0.y, fmaxf(v1.y, v2.y)), fmaxf(v0.z, fmaxf(v1.z, v2.z)));
  float3 c   = (lo3 + hi3) * 0.5f;
  float3 ext = bmax - bmin;
  const float scale = 2097151.0f;  // 2^21 - 1
  float fx = ext.x > 0.f ? (c.x - bmin.x) / ext.x : 0.f;
  float fy = ext.y > 0.f ? (c.y - bmin.y) / ext.y : 0.f;
  float fz = ext.z > 0.f ? (c.z - bmin.z) / ext.z : 0.f;
  uint32_t ix = (uint32_t)fminf(fmaxf(fx * scale, 0.f), scale);
  uint32_t iy = (uint32_t)fminf(fmaxf(fy * scale, 0.f), scale);
  uint32_t iz = (uint32_t)fminf(fmaxf(fz * scale, 0.f), scale);
  keys[gid]   = (expandBits21(ix) << 2) | (expandBits21(iy) << 1) | expandBits21(iz);
  vals[gid]   = gid;
}

// ---- 2b. binned-SAH path codes (VKRT_BUILD_SAH_GPU) ---------------------------------------------------------
// The builder the reference asks its driver for is PREFER_FAST_TRACE (src/accelstruct.cpp:125-126).  On the GPU that is a top-down
// binned surface-area-heuristic build, breadth first, parallel over triangles:
//   per level   bin every triangle of every open node into 32 bins per axis (global atomics on ordered-uint boxes),
//               one warp per node sweeps the bins (lane = bin, shuffle scans) and picks the cheapest axis / plane,
//               every triangle moves to its child and appends one bit to its path code.
// Nodes with <= kSahStop triangles are closed.  The key of a triangle is its path code, left aligned in 30 bits, followed by a
// 33-bit Morton code of its centroid inside the closed node: path codes are prefix free, so the binary radix tree Karras' algorithm
// builds over the sorted keys has exactly the SAH topology on top and a Morton order inside the small closed nodes -- and the rest
// of the LBVH pipeline (radix sort, hierarchy, refit with SAH leaf collapse, emit) runs unchanged.
struct Box
{
  float lo[3], hi[3];
};
constexpr int      kSahBins     = 32;
constexpr int      kSahPathBits = 30;
struct SahSlot
{
  uint32_t cb[6];      // centroid bounds, ordered uints: lo xyz, hi xyz
  uint32_t count;
  uint32_t childBase;  // slots of the next level (left, right) when split
  int32_t  axis, bin;  // axis < 0: closed
  uint32_t finalIdx;   // closed: index into the closed-node boxes
  int32_t  binIdx;     // open (count > stop): its block of 3 x 32 bins this level; -1: will be closed without a look at the bins
};
struct SahBin
{
  uint32_t count;
  uint32_t lo[3], hi[3];  // ordered uints
};
struct SahLevel
{
  uint32_t numSlots, numOpen;
};

__global__ void sahRootKernel(SahSlot* slots, const uint32_t* __restrict__ boundsOrdered, uint32_t numTris, SahLevel* levels, uint32_t* finalCount,
                              uint32_t stop)
{
  SahSlot s;
  for(int k = 0; k < 6; k++) s.cb[k] = boundsOrdered[k];
  s.count = numTris; s.childBase = 0; s.axis = -1; s.bin = 0; s.finalIdx = 0;
  s.binIdx = numTris > stop ? 0 : -1;
  slots[0] = s;
  for(int l = 0; l <= kSahPathBits + 1; l++) { levels[l].numSlots = 0; levels[l].numOpen = 0; }
  levels[0].numSlots = 1;
  levels[0].numOpen  = numTris > stop ? 1u : 0u;
  *finalCount = 0;
}

// bins of the open nodes of this level: count 0, boxes inverted
__global__ void sahBinClearKernel(SahBin* bins, const SahLevel* __restrict__ level)
{
  const size_t n = (size_t)level->numOpen * 3 * kSahBins;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    SahBin b;
    b.count = 0;
    b.lo[0] = b.lo[1] = b.lo[2] = 0xffffffffu;
    b.hi[0] = b.hi[1] = b.hi[2] = 0u;
    bins[i] = b;
  }
}

__device__ __forceinline__ void triBoundsCentroid(const float4* __restrict__ wtri, uint32_t i, float (&lo)[3], float (&hi)[3], float (&c)[3])
{
  const float4 a = wtri[3ull * i], b = wtri[3ull * i + 1], d = wtri[3ull * i + 2];
  lo[0] = fminf(a.x, fminf(b.x, d.x)); lo[1] = fminf(a.y, fminf(b.y, d.y)); lo[2] = fminf(a.z, fminf(b.z, d.z));
  hi[0] = fmaxf(a.x, fmaxf(b.x, d.x)); hi[1] = fmaxf(a.y, fmaxf(b.y, d.y)); hi[2] = fmaxf(a.z, fmaxf(b.z, d.z));
  for(int k = 0; k < 3; k++) c[k] = (lo[k] + hi[k]) * 0.5f;
}
__device__ __forceinline__ int sahBinOf(float c, float cmin, float cmax)
{
  const float scale = (float)kSahBins / (cmax - cmin);
  return min(kSahBins - 1, max(0, (int)((c - cmin) * scale)));
}

__global__ void sahBinKernel(const float4* __restrict__ wtri, uint32_t numTris, const int32_t* __restrict__ triSlot, const SahSlot* __restrict__ slots,
                             SahBin* __restrict__ bins)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= numTris)
    return;
  const int32_t sidx = triSlot[i];
  if(sidx < 0)
    return;
  const SahSlot& S = slots[sidx];
  if(S.binIdx < 0)
    return;  // will be closed: no split to evaluate
  float lo[3], hi[3], c[3];
  triBoundsCentroid(wtri, i, lo, hi, c);
  const uint32_t olo[3] = {orderedFromFloat(lo[0]), orderedFromFloat(lo[1]), orderedFromFloat(lo[2])};
  const uint32_t ohi[3] = {orderedFromFloat(hi[0]), orderedFromFloat(hi[1]), orderedFromFloat(hi[2])};
  for(int a = 0; a < 3; a++)
  {
    const float cmin = floatFromOrdered(S.cb[a]), cmax = floatFromOrdered(S.cb[3 + a]);
    if(!(cmax > cmin))
      continue;
    // warp-aggregated: neighbouring triangles of a mesh mostly fall into the same bin of the same node, and near the root every
    // triangle of the scene hits the same 96 bins -- lanes with the same target combine (match + reduce) and one of them does the atomics
    const uint32_t key = (uint32_t)(((size_t)S.binIdx * 3 + a) * kSahBins + sahBinOf(c[a], cmin, cmax));
    const unsigned m   = __match_any_sync(__activemask(), key);
    const uint32_t l0 = __reduce_min_sync(m, olo[0]), l1 = __reduce_min_sync(m, olo[1]), l2 = __reduce_min_sync(m, olo[2]);
    const uint32_t h0 = __reduce_max_sync(m, ohi[0]), h1 = __reduce_max_sync(m, ohi[1]), h2 = __reduce_max_sync(m, ohi[2]);
    if((int)(threadIdx.x & 31u) == __ffs(m) - 1)
    {
      SahBin* b = bins + key;
      atomicAdd(&b->count, (uint32_t)__popc(m));
      atomicMin(&b->lo[0], l0); atomicMin(&b->lo[1], l1); atomicMin(&b->lo[2], l2);
      atomicMax(&b->hi[0], h0); atomicMax(&b->hi[1], h1); atomicMax(&b->hi[2], h2);
    }
  }
}

__device__ __forceinline__ float sahArea(const float (&lo)[3], const float (&hi)[3])
{
  const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
  return 2.0f * (x * y + y * z + z * x);
}

// one warp per node of this level, lane = bin
__global__ void sahSplitKernel(SahSlot* __restrict__ slots, SahLevel* __restrict__ levels, int L, const SahBin* __restrict__ bins,
                               SahSlot* __restrict__ next, Box* __restrict__ finals, uint32_t* __restrict__ finalCount, uint32_t stop, uint32_t maxSlots)
{
  const uint32_t sidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int      lane = threadIdx.x & 31;
  if(sidx >= levels[L].numSlots)
    return;
  SahSlot  S = slots[sidx];
  float    bestCost = 3e38f;
  int      bestAxis = -1, bestBin = -1;
  uint32_t bestL = 0;
  float    nodeArea = 0.f;
  if(S.binIdx >= 0)
  {
    for(int a = 0; a < 3; a++)
    {
      const float cmin = floatFromOrdered(S.cb[a]), cmax = floatFromOrdered(S.cb[3 + a]);
      if(!(cmax > cmin))
        continue;
      const SahBin b = bins[((size_t)S.binIdx * 3 + a) * kSahBins + lane];
      // inclusive prefix (left) and suffix (right) unions over the lanes
      float    llo[3], lhi[3], rlo[3], rhi[3];
      uint32_t lc = b.count, rc = b.count;
      for(int k = 0; k < 3; k++)
      {
        llo[k] = rlo[k] = b.count ? floatFromOrdered(b.lo[k]) : 3e38f;
        lhi[k] = rhi[k] = b.count ? floatFromOrdered(b.hi[k]) : -3e38f;
      }
#pragma unroll
      for(int o = 1; o < 32; o <<= 1)
      {
        const uint32_t c1 = __shfl_up_sync(0xffffffffu, lc, o), c2 = __shfl_down_sync(0xffffffffu, rc, o);
        float t[12];
        for(int k = 0; k < 3; k++)
        {
          t[k]     = __shfl_up_sync(0xffffffffu, llo[k], o);
          t[3 + k] = __shfl_up_sync(0xffffffffu, lhi[k], o);
          t[6 + k] = __shfl_down_sync(0xffffffffu, rlo[k], o);
          t[9 + k] = __shfl_down_sync(0xffffffffu, rhi[k], o);
        }
        if(lane >= o)
        {
          lc += c1;
          for(int k = 0; k < 3; k++) { llo[k] = fminf(llo[k], t[k]); lhi[k] = fmaxf(lhi[k], t[3 + k]); }
        }
        if(lane + o < 32)
        {
          rc += c2;
          for(int k = 0; k < 3; k++) { rlo[k] = fminf(rlo[k], t[6 + k]); rhi[k] = fmaxf(rhi[k], t[9 + k]); }
        }
      }
      // split after bin `lane`: left = prefix(lane), right = suffix(lane + 1)
      const uint32_t rcn = __shfl_down_sync(0xffffffffu, rc, 1);
      float          nlo[3], nhi[3];
      for(int k = 0; k < 3; k++)
      {
        nlo[k] = __shfl_down_sync(0xffffffffu, rlo[k], 1);
        nhi[k] = __shfl_down_sync(0xffffffffu, rhi[k], 1);
      }
      float cost = 3e38f;
      if(lane < 31 && lc > 0 && rcn > 0)
        cost = sahArea(llo, lhi) * (float)lc + sahArea(nlo, nhi) * (float)rcn;
      // the node's own box: prefix of the last lane
      {
        float flo[3], fhi[3];
        for(int k = 0; k < 3; k++) { flo[k] = __shfl_sync(0xffffffffu, llo[k], 31); fhi[k] = __shfl_sync(0xffffffffu, lhi[k], 31); }
        nodeArea = sahArea(flo, fhi);
      }
      // warp argmin, ties to the lower bin; axes are visited in order, strict improvement only (as the host builder)
      float    c  = cost;
      int      bi = lane;
      uint32_t cl = lc;
#pragma unroll
      for(int o = 16; o > 0; o >>= 1)
      {
        const float    oc = __shfl_xor_sync(0xffffffffu, c, o);
        const int      ob = __shfl_xor_sync(0xffffffffu, bi, o);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, cl, o);
        if(oc < c || (oc == c && ob < bi)) { c = oc; bi = ob; cl = ol; }
      }
      if(c < bestCost) { bestCost = c; bestAxis = a; bestBin = bi; bestL = cl; }
    }
  }
  if(lane != 0)
    return;
  // close the node when it is small, cannot be split, is as deep as the path code, or a leaf would be cheaper anyway
  bool split = bestAxis >= 0;
  if(split && S.count <= 4u && nodeArea * (float)S.count <= nodeArea + bestCost)
    split = false;
  uint32_t base = 0;
  if(split)
  {
    base = atomicAdd(&levels[L + 1].numSlots, 2u);
    if(base + 2u > maxSlots)
      split = false;  // cannot happen for counts that add up (children of open nodes only); keeps a corrupt input from writing out of bounds
  }
  if(split)
  {
    S.childBase = base; S.axis = bestAxis; S.bin = bestBin;
    const bool     nextIsLast = (L + 1) >= kSahPathBits;
    const uint32_t cnt[2]     = {bestL, S.count - bestL};
    for(int k = 0; k < 2; k++)
    {
      SahSlot ch;
      ch.cb[0] = ch.cb[1] = ch.cb[2] = 0xffffffffu;
      ch.cb[3] = ch.cb[4] = ch.cb[5] = 0u;
      ch.count = cnt[k]; ch.childBase = 0; ch.axis = -1; ch.bin = 0; ch.finalIdx = 0;
      ch.binIdx = (cnt[k] > stop && !nextIsLast) ? (int32_t)atomicAdd(&levels[L + 1].numOpen, 1u) : -1;
      next[base + k] = ch;
    }
  }
  else
  {
    S.axis     = -1;
    S.finalIdx = atomicAdd(finalCount, 1u);
    Box f;
    for(int k = 0; k < 3; k++) { f.lo[k] = floatFromOrdered(S.cb[k]); f.hi[k] = floatFromOrdered(S.cb[3 + k]); }
    finals[S.finalIdx] = f;
  }
  slots[sidx] = S;
}

__global__ void sahApplyKernel(const float4* __restrict__ wtri, uint32_t numTris, int32_t* __restrict__ triSlot, uint32_t* __restrict__ triCode,
                               uint32_t* __restrict__ triDepth, uint32_t* __restrict__ triFinal, const SahSlot* __restrict__ slots,
                               SahSlot* __restrict__ next)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= numTris)
    return;
  const int32_t sidx = triSlot[i];
  if(sidx < 0)
    return;
  const SahSlot& S = slots[sidx];
  if(S.axis < 0)
  {
    triFinal[i] = S.finalIdx;
    triSlot[i]  = -1;
    return;
  }
  float lo[3], hi[3], c[3];
  triBoundsCentroid(wtri, i, lo, hi, c);
  const int      k    = sahBinOf(c[S.axis], floatFromOrdered(S.cb[S.axis]), floatFromOrdered(S.cb[3 + S.axis]));
  const uint32_t side = k > S.bin ? 1u : 0u;
  const uint32_t ch   = S.childBase + side;
  triSlot[i]  = (int32_t)ch;
  triCode[i]  = (triCode[i] << 1) | side;
  triDepth[i] = triDepth[i] + 1u;
  // centroid bounds of the child, warp-aggregated like the bins
  const unsigned m = __match_any_sync(__activemask(), ch);
  uint32_t       lo3[3], hi3[3];
  for(int a = 0; a < 3; a++)
  {
    const uint32_t oc = orderedFromFloat(c[a]);
    lo3[a] = __reduce_min_sync(m, oc);
    hi3[a] = __reduce_max_sync(m, oc);
  }
  if((int)(threadIdx.x & 31u) == __ffs(m) - 1)
    for(int a = 0; a < 3; a++)
    {
      atomicMin(&next[ch].cb[a], lo3[a]);
      atomicMax(&next[ch].cb[3 + a], hi3[a]);
    }
}

__device__ __forceinline__ uint64_t expandBits11(uint32_t v)  // bit b of v -> bit 3b
{
  uint64_t r = 0;
#pragma unroll
  for(int b = 0; b < 11; b++) r |= (uint64_t)((v >> b) & 1u) << (3 * b);
  return r;
}
__global__ void sahKeyKernel(const float4* __restrict__ wtri, uint32_t numTris, const uint32_t* __restrict__ triCode, const uint32_t* __restrict__ triDepth,
                             const uint32_t* __restrict__ triFinal, const Box* __restrict__ finals, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= numTris)
    return;
  float lo[3], hi[3], c[3];
  triBoundsCentroid(wtri, i, lo, hi, c);
  const Box      f = finals[triFinal[i]];
  uint32_t       q[3];
  for(int k = 0; k < 3; k++)
  {
    const float ext = f.hi[k] - f.lo[k];
    const float t   = ext > 0.f ? (c[k] - f.lo[k]) / ext : 0.f;
    q[k]            = (uint32_t)fminf(fmaxf(t * 2047.0f, 0.f), 2047.0f);
  }
  const uint64_t morton = (expandBits11(q[0]) << 2) | (expandBits11(q[1]) << 1) | expandBits11(q[2]);
  const uint32_t d      = triDepth[i];
  const uint64_t path   = (uint64_t)triCode[i] << (kSahPathBits - (int)d);
  keys[i] = (path << 33) | morton;
  vals[i] = i;
}

// ---- 4. hierarchy (Karras 2012) ----------------------------------------------------------
__device__ __forceinline__ int deltaKeys(const uint64_t* __restrict__ keys, int n, int i, int j)
{
  if(j < 0 || j >= n)
    return -1;
  uint64_t a = keys[i], b = keys[j];
  if(a == b)
    return 64 + __clz((uint32_t)i ^ (uint32_t)j);
  return __clzll((long long)(a ^ b));
}

// child codes during the build: >= 0 internal node index, < 0 : ~leafIndex
__global__ void hierarchyKernel(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children, int2* __restrict__ ranges,
                                int* __restrict__ parentInner, int* __restrict__ parentLeaf)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n - 1)
    return;
  int d     = (deltaKeys(keys, n, i, i + 1) - deltaKeys(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin  = deltaKeys(keys, n, i, i - d);
  int lmax  = 2;
  while(deltaKeys(keys, n, i, i + lmax * d) > dmin)
    lmax <<= 1;
  int l = 0;
  for(int t = lmax >> 1; t >= 1; t >>= 1)
    if(deltaKeys(keys, n, i, i + (l + t) * d) > dmin)
      l += t;
  int j     = i + l * d;
  int dnode = deltaKeys(keys, n, i, j);
  int s     = 0;
  int t     = l;
  do
  {
    t = (t + 1) >> 1;
    if(deltaKeys(keys, n, i, i + (s + t) * d) > dnode)
      s += t;
  } while(t > 1);
  int gamma = i + s * d + min(d, 0);
  int first = min(i, j), last = max(i, j);
  int left  = (first == gamma) ? ~gamma : gamma;
  int right = (last == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
  children[i] = make_int2(left, right);
  ranges[i]   = make_int2(first, last);
  if(left >= 0) parentInner[left] = i; else parentLeaf[~left] = i;
  if(right >= 0) parentInner[right] = i; else parentLeaf[~right] = i;
  if(i == 0)
    parentInner[0] = -1;
}

// ---- 5. leaves + refit ----------------------------------------------------------------------
// Boxes / costs written by other blocks during the bottom-up pass must not be read
// through a stale L1 line: load them with ld.global.cg.
__device__ __forceinline__ Box loadBoxCg(const Box* p)
{
  Box          b;
  const float* f = reinterpret_cast<const float*>(p);
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    b.lo[k] = __ldcg(f + k);
    b.hi[k] = __ldcg(f + 3 + k);
  }
  return b;
}
__device__ __forceinline__ float boxArea(const Box& b)
{
  float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
  return 2.0f * (ex * ey + ey * ez + ez * ex);
}

__global__ void refitKernel(const float4* __restrict__ wtri, const uint32_t* __restrict__ sortedVals, int n, float4* __restrict__ tris,
                            const int2* __restrict__ children, const int* __restrict__ parentInner, const int* __restrict__ parentLeaf,
                            Box* __restrict__ leafBox, Box* __restrict__ nodeBox, float* __restrict__ nodeCost, int* __restrict__ nodeDepth,
                            uint32_t* __restrict__ collapsed, const int2* __restrict__ ranges, int* __restrict__ flags, int collapseMax,
                            int* __restrict__ maxDepthOut, const uint64_t* __restrict__ groupKeys = nullptr)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j >= n)
    return;
  uint32_t gid = sortedVals[j];
  float4   a = wtri[3ull * gid], b = wtri[3ull * gid + 1], c = wtri[3ull * gid + 2];
  tris[3ull * j] = a; tris[3ull * j + 1] = b; tris[3ull * j + 2] = c;
  Box bx;
  bx.lo[0] = fminf(a.x, fminf(b.x, c.x)); bx.lo[1] = fminf(a.y, fminf(b.y, c.y)); bx.lo[2] = fminf(a.z, fminf(b.z, c.z));
  bx.hi[0] = fmaxf(a.x, fmaxf(b.x, c.x)); bx.hi[1] = fmaxf(a.y, fmaxf(b.y, c.y)); bx.hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
  leafBox[j] = bx;
  if(n == 1)
    return;
  // groupKeys (VKRT_BUILD_SAH_GPU): sorted path-code keys.  Triangles with the same path code are one closed node of the SAH
  // build; only inside such a group may a leaf be formed (with the host builder's rule: a leaf when splitting does not pay,
  // inner-node cost 1.0), across groups the SAH split stands.
  const float Ct = 1.0f, Ci = groupKeys ? 1.0f : 1.2f;
  int node = parentLeaf[j];
  __threadfence();
  while(node >= 0)
  {
    if(atomicAdd(&flags[node], 1) == 0)
      return;  // first arrival: the sibling subtree is not finished yet
    __threadfence();
    int2 ch = children[node];
    Box  l  = loadBoxCg(ch.x >= 0 ? &nodeBox[ch.x] : &leafBox[~ch.x]);
    Box  r  = loadBoxCg(ch.y >= 0 ? &nodeBox[ch.y] : &leafBox[~ch.y]);
    Box  u;
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      u.lo[k] = fminf(l.lo[k], r.lo[k]);
      u.hi[k] = fmaxf(l.hi[k], r.hi[k]);
    }
    nodeBox[node] = u;
    float cl = ch.x >= 0 ? __ldcg(&nodeCost[ch.x]) : Ct * boxArea(l);
    float cr = ch.y >= 0 ? __ldcg(&nodeCost[ch.y]) : Ct * boxArea(r);
    int   dl = ch.x >= 0 ? __ldcg(&nodeDepth[ch.x]) : 0;
    int   dr = ch.y >= 0 ? __ldcg(&nodeDepth[ch.y]) : 0;
    int2  rg    = ranges[node];
    int   count = rg.y - rg.x + 1;
    float A     = boxArea(u);
    float cInner = Ci * A + cl + cr;
    float cLeaf  = Ct * A * (float)count;
    bool  col    = collapseMax > 1 && count <= collapseMax && cLeaf <= cInner;
    if(groupKeys)
      col = col && ((groupKeys[rg.x] >> 33) == (groupKeys[rg.y] >> 33));
    collapsed[node] = col ? 1u : 0u;
    nodeCost[node]  = col ? cLeaf : cInner;
    int depth       = col ? 0 : (max(dl, dr) + 1);
    nodeDepth[node] = depth;
    if(parentInner[node] < 0)
      *maxDepthOut = depth;
    __threadfence();
    node = parentInner[node];
  }
}

// ---- 4b. PLOC topology (Meister & Bittner 2017): parallel locally-ordered clustering ----------------------
// An alternative to the Karras hierarchy with SAH-like quality.  The Morton-sorted leaves are the initial clusters; every
// round each cluster looks R places left and right for the neighbour whose union box has the smallest area, mutual nearest
// neighbours merge, and the survivors are compacted in place (order is kept, so Morton locality survives).  ~40 % of the
// clusters disappear per round.  The result is written in the conventions of hierarchyKernel (root = node 0, child codes,
// parent arrays, [first,last] leaf ranges) after the leaves have been renumbered in depth-first order, so that refitKernel /
// emitKernel and the multi-triangle leaf collapse run unchanged.
// search radius: 16 while there are many clusters; wider near the top of the tree, where every ray passes and the Morton
// order says little about which big clusters belong together
static int plocRadius(int c) { (void)c; return 16; }  // wider windows near the top were measured and traced worse (profiles/README.md)

__global__ void plocInitKernel(const float4* __restrict__ wtri, const uint32_t* __restrict__ sortedVals, int n, int* __restrict__ cid,
                               Box* __restrict__ cbox, int* __restrict__ csize)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j >= n)
    return;
  const uint32_t gid = sortedVals[j];
  const float4   a = wtri[3ull * gid], b = wtri[3ull * gid + 1], c = wtri[3ull * gid + 2];
  Box bx;
  bx.lo[0] = fminf(a.x, fminf(b.x, c.x)); bx.lo[1] = fminf(a.y, fminf(b.y, c.y)); bx.lo[2] = fminf(a.z, fminf(b.z, c.z));
  bx.hi[0] = fmaxf(a.x, fmaxf(b.x, c.x)); bx.hi[1] = fmaxf(a.y, fmaxf(b.y, c.y)); bx.hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
  cid[j]   = ~j;
  cbox[j]  = bx;
  csize[j] = 1;
}

__device__ __forceinline__ float unionArea(const Box& a, const Box& b)
{
  const float ex = fmaxf(a.hi[0], b.hi[0]) - fminf(a.lo[0], b.lo[0]);
  const float ey = fmaxf(a.hi[1], b.hi[1]) - fminf(a.lo[1], b.lo[1]);
  const float ez = fmaxf(a.hi[2], b.hi[2]) - fminf(a.lo[2], b.lo[2]);
  return ex * ey + ey * ez + ez * ex;
}

// nn[i] = argmin over the window of (union area, j): ascending scan + strict '<' makes the globally best pair mutual, so
// every round merges at least one pair.  forcePair (slow progress, e.g. thousands of identical boxes): neighbours 2k, 2k+1.
__global__ void plocNearestKernel(const Box* __restrict__ cbox, int c, int* __restrict__ nn, int forcePair, int radius)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= c)
    return;
  if(forcePair)
  {
    const int j = i ^ 1;
    nn[i]       = j < c ? j : -1;
    return;
  }
  const Box me   = cbox[i];
  int       best = -1;
  float     bd   = FLT_MAX;
  const int j0 = max(0, i - radius), j1 = min(c - 1, i + radius);
  for(int j = j0; j <= j1; j++)
  {
    if(j == i)
      continue;
    const float d = unionArea(me, cbox[j]);
    if(d < bd || best < 0)
    {
      bd   = d;
      best = j;
    }
  }
  nn[i] = best;
}

__global__ void plocMergeKernel(const int* __restrict__ cid, const Box* __restrict__ cbox, const int* __restrict__ csize,
                                const int* __restrict__ nn, int c, int n, int* __restrict__ nodeCounter, int2* __restrict__ children,
                                int* __restrict__ parentInner, int* __restrict__ parentLeaf, int* __restrict__ nodeSize,
                                int* __restrict__ cidOut, Box* __restrict__ boxOut, int* __restrict__ sizeOut, uint32_t* __restrict__ valid)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= c)
    return;
  const int  j      = nn[i];
  const bool mutual = j >= 0 && nn[j] == i;
  if(!mutual)
  {
    valid[i]   = 1u;
    cidOut[i]  = cid[i];
    boxOut[i]  = cbox[i];
    sizeOut[i] = csize[i];
    return;
  }
  if(i > j)
  {
    valid[i] = 0u;
    return;
  }
  const int idx = (n - 2) - atomicAdd(nodeCounter, 1);  // the last merge becomes node 0, the root
  const int a = cid[i], b = cid[j];
  children[idx] = make_int2(a, b);
  if(a >= 0) parentInner[a] = idx; else parentLeaf[~a] = idx;
  if(b >= 0) parentInner[b] = idx; else parentLeaf[~b] = idx;
  const Box x = cbox[i], y = cbox[j];
  Box       u;
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    u.lo[k] = fminf(x.lo[k], y.lo[k]);
    u.hi[k] = fmaxf(x.hi[k], y.hi[k]);
  }
  const int sz  = csize[i] + csize[j];
  nodeSize[idx] = sz;
  valid[i]      = 1u;
  cidOut[i]     = idx;
  boxOut[i]     = u;
  sizeOut[i]    = sz;
}

__global__ void plocCompactKernel(const uint32_t* __restrict__ valid, const uint32_t* __restrict__ pos, int c, const int* __restrict__ cidIn,
                                  const Box* __restrict__ boxIn, const int* __restrict__ sizeIn, int* __restrict__ cid, Box* __restrict__ cbox,
                                  int* __restrict__ csize, int* __restrict__ newCount)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= c)
    return;
  if(valid[i])
  {
    const uint32_t p = pos[i];
    cid[p]   = cidIn[i];
    cbox[p]  = boxIn[i];
    csize[p] = sizeIn[i];
  }
  if(i == c - 1)
    *newCount = (int)(pos[i] + valid[i]);
}

// depth-first position of a leaf / first leaf of an inner node: the sizes of the left siblings on the way up
__device__ __forceinline__ int plocOffset(int code, int parent, const int2* __restrict__ children, const int* __restrict__ parentInner,
                                          const int* __restrict__ nodeSize)
{
  int off = 0, cur = code, p = parent;
  while(p >= 0)
  {
    const int2 ch = children[p];
    if(ch.y == cur)
      off += ch.x >= 0 ? nodeSize[ch.x] : 1;
    cur = p;
    p   = parentInner[p];
  }
  return off;
}
__global__ void plocOrderKernel(int n, const int2* __restrict__ children, const int* __restrict__ parentInner, const int* __restrict__ parentLeaf,
                                const int* __restrict__ nodeSize, int* __restrict__ leafPos, int2* __restrict__ ranges)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n)
    leafPos[i] = plocOffset(~i, parentLeaf[i], children, parentInner, nodeSize);
  if(i < n - 1)
  {
    const int off = plocOffset(i, parentInner[i], children, parentInner, nodeSize);
    ranges[i]     = make_int2(off, off + nodeSize[i] - 1);
  }
}
__global__ void plocRemapKernel(int n, int2* __restrict__ children, const int* __restrict__ leafPos, const uint32_t* __restrict__ sortedVals,
                                uint32_t* __restrict__ sortedOut, const int* __restrict__ parentLeafOld, int* __restrict__ parentLeafNew)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n)
  {
    const int p      = leafPos[i];
    sortedOut[p]     = sortedVals[i];
    parentLeafNew[p] = parentLeafOld[i];
  }
  if(i < n - 1)
  {
    int2 ch = children[i];
    if(ch.x < 0) ch.x = ~leafPos[~ch.x];
    if(ch.y < 0) ch.y = ~leafPos[~ch.y];
    children[i] = ch;
  }
}

// ---- 6. emit ---------------------------------------------------------------------------
__global__ void emitKernel(int n, const int2* __restrict__ children, const int2* __restrict__ ranges, const Box* __restrict__ leafBox,
                           const Box* __restrict__ nodeBox, const uint32_t* __restrict__ collapsed, float4* __restrict__ nodes)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= max(n - 1, 1))
    return;
  Box      b0, b1;
  uint32_t c0, c1;
  if(n == 1)
  {
    // a single triangle: both children are the same leaf (the duplicate candidate
    // loses the tie-break, so it is harmless)
    b0 = leafBox[0];
    b1 = b0;
    c0 = kLeafBit | 0u;
    c1 = c0;
  }
  else
  {
    int2 ch = children[i];
    auto enc = [&](int c, Box& bx) -> uint32_t {
      if(c < 0)
      {
        bx = leafBox[~c];
        return kLeafBit | (uint32_t)(~c);
      }
      bx = nodeBox[c];
      if(collapsed[c])
      {
        int2 rg = ranges[c];
        return kLeafBit | ((uint32_t)(rg.y - rg.x) << kLeafCntSh) | (uint32_t)rg.x;
      }
      return (uint32_t)c;
    };
    c0 = enc(ch.x, b0);
    c1 = enc(ch.y, b1);
  }
  nodes[4ull * i + 0] = f4(b0.lo[0], b0.hi[0], b0.lo[1], b0.hi[1]);
  nodes[4ull * i + 1] = f4(b1.lo[0], b1.hi[0], b1.lo[1], b1.hi[1]);
  nodes[4ull * i + 2] = f4(b0.lo[2], b0.hi[2], b1.lo[2], b1.hi[2]);
  nodes[4ull * i + 3] = f4(__uint_as_float(c0), __uint_as_float(c1), 0.f, 0.f);
}

// numTris == 0: the trace kernels never traverse (they test SceneView::numTris); the node
// is only written so that the buffer is defined.
__global__ void emptySceneKernel(float4* nodes)
{
  nodes[0] = f4(0.f, 0.f, 0.f, 0.f);
  nodes[1] = f4(0.f, 0.f, 0.f, 0.f);
  nodes[2] = f4(0.f, 0.f, 0.f, 0.f);
  nodes[3] = f4(__uint_as_float(kLeafBit), __uint_as_float(kLeafBit), 0.f, 0.f);
}

#define CK(x)                                                                                                          \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e_ = (x);                                                                                              \
    if(e_ != cudaSuccess)                                                                                              \
    {                                                                                                                  \
      err = std::string(#x) + ": " + cudaGetErrorString(e_);                                                           \
      goto fail;                                                                                                       \
    }                                                                                                                  \
  } while(0)
#define TCK(x) CK(x)
// Build temporaries come from CUDA's stream-ordered pool (the context keeps up to 1 GB of it warm, vkrt_create): a build is ~40
// allocations, and plain cudaMalloc / cudaFree (synchronising, tens of microseconds to milliseconds each) cost more than the kernels.
template <class T>
static cudaError_t tmpAlloc(T** p, size_t bytes, cudaStream_t s)
{
  return cudaMallocAsync(reinterpret_cast<void**>(p), bytes ? bytes : 16, s);
}
static void tmpFree(void* p, cudaStream_t s)
{
  if(p)
    cudaFreeAsync(p, s);
}


// 32-byte nodes (VKRT_NODE_FORMAT=q32, and the box encoding of the 4-wide tree): a node that fits one 256-bit load.  On an
// L2-resident BVH it is no faster than the 64-byte node (the step is latency bound, profiles/README.md); it halves the
// footprint of an HBM-resident one.  The two child boxes are stored on a 16-bit grid
// G(p) = (p - gridOrigin) * gridInv laid over the root box with an 8-cell margin:
//   word 0..2: child 0  x,y,z  (lo | hi << 16)      word 3..5: child 1      word 6,7: child references (as in the fp32 node)
// lo is rounded down and hi up after padding by 1.0625 cells: the trace kernel evaluates the slab planes as
// fma(2^23 + q, 1/d_g, t0 - (o_g + 2^23) / d_g), whose rounding moves a plane by at most 0.5 (o_g + 2^23 rounds to a whole
// cell) + 0.504 (the product) + 2^-7 (evaluating G in fp32) cells, so the padded box always contains the true one and the
// box test stays a conservative filter: hits are decided by the triangle test alone, results are identical to the fp32 nodes.
__global__ void quantizeNodesKernel(const float4* __restrict__ nodes, uint32_t numNodes, float3 origin, float3 inv, uint4* __restrict__ q)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= numNodes)
    return;
  const float4 n0 = nodes[4ull * i], n1 = nodes[4ull * i + 1], n2 = nodes[4ull * i + 2], n3 = nodes[4ull * i + 3];
  auto lo = [](float v, float o, float s) { return (uint32_t)min(65535, max(0, (int)floorf((v - o) * s - 1.0625f))); };
  auto hi = [](float v, float o, float s) { return (uint32_t)min(65535, max(0, (int)ceilf((v - o) * s + 1.0625f))); };
  uint4 a, b;
  a.x = lo(n0.x, origin.x, inv.x) | (hi(n0.y, origin.x, inv.x) << 16);
  a.y = lo(n0.z, origin.y, inv.y) | (hi(n0.w, origin.y, inv.y) << 16);
  a.z = lo(n2.x, origin.z, inv.z) | (hi(n2.y, origin.z, inv.z) << 16);
  a.w = lo(n1.x, origin.x, inv.x) | (hi(n1.y, origin.x, inv.x) << 16);
  b.x = lo(n1.z, origin.y, inv.y) | (hi(n1.w, origin.y, inv.y) << 16);
  b.y = lo(n2.z, origin.z, inv.z) | (hi(n2.w, origin.z, inv.z) << 16);
  b.z = __float_as_uint(n3.x);
  b.w = __float_as_uint(n3.y);
  q[2ull * i]     = a;
  q[2ull * i + 1] = b;
}

// Centre / half-extent copy of the fp32 nodes, the layout the default traversal kernel reads:
//   m0 = (c0.cx, c0.ex, c0.cy, c0.ey)  m1 = (c1.cx, c1.ex, c1.cy, c1.ey)  m2 = (c0.cz, c0.ez, c1.cz, c1.ez)  m3 = references
// With it the slab test is t_near/far = fma(-/+e, |1/d|, fma(c, 1/d, -o/d)): nine FMA-pipe instructions per box and no
// per-axis min/max on the (half-rate, saturated) ALU pipe.  e is padded so that [c-e, c+e] contains the fp32 box and the
// rounding of c/d - o/d (a plane displacement of at most 2^-24 of the coordinate magnitude) can never cut a box short.
__global__ void centerNodesKernel(const float4* __restrict__ nodes, uint32_t numNodes, float4* __restrict__ out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= numNodes)
    return;
  const float4 n0 = nodes[4ull * i], n1 = nodes[4ull * i + 1], n2 = nodes[4ull * i + 2], n3 = nodes[4ull * i + 3];
  auto ce = [](float lo, float hi, float& c, float& e) {
    c = 0.5f * lo + 0.5f * hi;
    e = fmaxf(hi - c, c - lo);
    e = (e + (fabsf(c) + e) * 4.8e-7f) * 1.000001f;
  };
  float4 m0, m1, m2;
  ce(n0.x, n0.y, m0.x, m0.y); ce(n0.z, n0.w, m0.z, m0.w);
  ce(n1.x, n1.y, m1.x, m1.y); ce(n1.z, n1.w, m1.z, m1.w);
  ce(n2.x, n2.y, m2.x, m2.y); ce(n2.z, n2.w, m2.z, m2.w);
  out[4ull * i] = m0; out[4ull * i + 1] = m1; out[4ull * i + 2] = m2; out[4ull * i + 3] = n3;
}

bool quantizeBvh(cudaStream_t stream, BvhResult& bvh, std::string& err, uint64_t* launches)
{
  float          root[16];
  const uint32_t n = std::max<uint32_t>(bvh.numNodes, 1u);
  cudaFree(bvh.qnodes);
  bvh.qnodes = nullptr;
  CK(cudaMalloc(&bvh.qnodes, 32ull * n));
  CK(cudaMemcpyAsync(root, bvh.nodes, sizeof(root), cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  {
    // root box = union of the root's child boxes (layout of emitKernel)
    const float lo[3] = {std::min(root[0], root[4]), std::min(root[2], root[6]), std::min(root[8], root[10])};
    const float hi[3] = {std::max(root[1], root[5]), std::max(root[3], root[7]), std::max(root[9], root[11])};
    float       maxExt = 0.f;
    for(int a = 0; a < 3; a++)
      maxExt = std::max(maxExt, hi[a] - lo[a]);
    if(!(maxExt > 0.f) || !std::isfinite(maxExt))
      maxExt = 1.f;
    for(int a = 0; a < 3; a++)
    {
      // a flat axis still gets cells of a useful size, so that ray origins stay within range of the 2^23 bias
      const float ext  = std::max(std::isfinite(hi[a] - lo[a]) ? hi[a] - lo[a] : 0.f, 1e-3f * maxExt);
      const float cell = ext / 65500.0f;
      bvh.gridOrigin[a] = (std::isfinite(lo[a]) ? lo[a] : 0.f) - 8.0f * cell;
      bvh.gridInv[a]    = 1.0f / cell;
    }
  }
  quantizeNodesKernel<<<(n + 255) / 256, 256, 0, stream>>>(bvh.nodes, n, make_float3(bvh.gridOrigin[0], bvh.gridOrigin[1], bvh.gridOrigin[2]),
                                                           make_float3(bvh.gridInv[0], bvh.gridInv[1], bvh.gridInv[2]), bvh.qnodes);
  cudaFree(bvh.cnodes);
  bvh.cnodes = nullptr;
  CK(cudaMalloc(&bvh.cnodes, 64ull * n));
  centerNodesKernel<<<(n + 255) / 256, 256, 0, stream>>>(bvh.nodes, n, bvh.cnodes);
  if(launches) (*launches) += 2;
  CK(cudaGetLastError());
  return true;
fail:
  cudaFree(bvh.qnodes); cudaFree(bvh.cnodes);
  bvh.qnodes = nullptr;
  bvh.cnodes = nullptr;
  return false;
}

// ---------------------------------------------------------------------------------------------
// BVH4: the binary tree collapsed into 4-wide nodes, boxes on the same 16-bit grid (64 bytes per node):
//   word 3k..3k+2: child k  x,y,z (lo | hi << 16), k = 0..3      word 12..15: child references
// A reference is a BVH4 node index or a leaf code (kLeafBit, as in the binary tree); empty slots carry an inverted box
// (lo 65535, hi 0) that no ray can enter and repeat the reference of child 0.  One traversal step tests four boxes behind two 256-bit
// loads, and a ray needs about half as many dependent steps as in the binary tree.
// Built level by level from the root: each 4-wide node starts from the two children of its binary node and
// repeatedly opens the inner child with the largest surface area until it has four children.
constexpr uint32_t kEmptyRef = 0xffffffffu;  // also has kLeafBit set: count field 7, never produced by emitKernel for a real leaf first index

struct WideCand
{
  float    lo[3], hi[3];
  uint32_t ref;
};

__global__ void collapseWideKernel(const float4* __restrict__ nodes, uint32_t* __restrict__ src, uint32_t begin, uint32_t end,
                                   uint32_t* __restrict__ count, float3 origin, float3 inv, uint4* __restrict__ wide)
{
  const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= end)
    return;
  WideCand c[4];
  int      n = 0;
  auto open = [&](uint32_t node, int slot) {
    const float4 n0 = nodes[4ull * node], n1 = nodes[4ull * node + 1], n2 = nodes[4ull * node + 2], n3 = nodes[4ull * node + 3];
    WideCand a, b;
    a.lo[0] = n0.x; a.hi[0] = n0.y; a.lo[1] = n0.z; a.hi[1] = n0.w; a.lo[2] = n2.x; a.hi[2] = n2.y; a.ref = __float_as_uint(n3.x);
    b.lo[0] = n1.x; b.hi[0] = n1.y; b.lo[1] = n1.z; b.hi[1] = n1.w; b.lo[2] = n2.z; b.hi[2] = n2.w; b.ref = __float_as_uint(n3.y);
    c[slot] = a;
    c[n++]  = b;
  };
  n = 1;  // slot 0 is filled by the first open()
  open(src[i], 0);
  while(n < 4)
  {
    int   pick = -1;
    float best = -1.f;
    for(int k = 0; k < n; k++)
    {
      if(c[k].ref & kLeafBit)
        continue;
      const float dx = c[k].hi[0] - c[k].lo[0], dy = c[k].hi[1] - c[k].lo[1], dz = c[k].hi[2] - c[k].lo[2];
      const float area = dx * dy + dy * dz + dz * dx;
      if(area > best)
      {
        best = area;
        pick = k;
      }
    }
    if(pick < 0)
      break;
    open(c[pick].ref, pick);
  }
  // inner children become 4-wide nodes of the next level
  int inner = 0;
  for(int k = 0; k < n; k++)
    inner += (c[k].ref & kLeafBit) ? 0 : 1;
  uint32_t base = inner ? atomicAdd(count, (uint32_t)inner) : 0u;
  uint32_t w[16];
  auto qlo = [](float v, float o, float s) { return (uint32_t)min(65535, max(0, (int)floorf((v - o) * s - 1.0625f))); };
  auto qhi = [](float v, float o, float s) { return (uint32_t)min(65535, max(0, (int)ceilf((v - o) * s + 1.0625f))); };
  for(int k = 0; k < 4; k++)
  {
    if(k < n)
    {
      w[3 * k + 0] = qlo(c[k].lo[0], origin.x, inv.x) | (qhi(c[k].hi[0], origin.x, inv.x) << 16);
      w[3 * k + 1] = qlo(c[k].lo[1], origin.y, inv.y) | (qhi(c[k].hi[1], origin.y, inv.y) << 16);
      w[3 * k + 2] = qlo(c[k].lo[2], origin.z, inv.z) | (qhi(c[k].hi[2], origin.z, inv.z) << 16);
      if(c[k].ref & kLeafBit)
        w[12 + k] = c[k].ref;
      else
      {
        src[base]  = c[k].ref;
        w[12 + k]  = base++;
      }
    }
    else
    {
      w[3 * k + 0] = w[3 * k + 1] = w[3 * k + 2] = 0x0000ffffu;  // lo 65535, hi 0: cannot be entered
      w[12 + k]    = w[12];  // ...and if rounding at astronomic distances ever let it through, it only repeats child 0
    }
  }
  uint4* out = wide + 4ull * i;
  out[0] = make_uint4(w[0], w[1], w[2], w[3]);
  out[1] = make_uint4(w[4], w[5], w[6], w[7]);
  out[2] = make_uint4(w[8], w[9], w[10], w[11]);
  out[3] = make_uint4(w[12], w[13], w[14], w[15]);
}

// needs quantizeBvh (grid) first
bool collapseBvhWide(cudaStream_t stream, BvhResult& bvh, std::string& err, uint64_t* launches)
{
  uint32_t *src = nullptr, *count = nullptr;
  uint32_t  begin = 0, end = 1, levels = 0;
  const uint32_t n = std::max<uint32_t>(bvh.numNodes, 1u);
  const float3   origin = make_float3(bvh.gridOrigin[0], bvh.gridOrigin[1], bvh.gridOrigin[2]);
  const float3   inv    = make_float3(bvh.gridInv[0], bvh.gridInv[1], bvh.gridInv[2]);
  cudaFree(bvh.wnodes);
  bvh.wnodes = nullptr;
  CK(cudaMalloc(&bvh.wnodes, 64ull * n));
  CK(tmpAlloc(&src, 4ull * n, stream));
  CK(tmpAlloc(&count, 4, stream));
  CK(cudaMemsetAsync(src, 0, 4, stream));  // the root of the binary tree is node 0
  {
    const uint32_t one = 1;
    CK(cudaMemcpyAsync(count, &one, 4, cudaMemcpyHostToDevice, stream));
  }
  while(begin < end)
  {
    collapseWideKernel<<<(end - begin + 127) / 128, 128, 0, stream>>>(bvh.nodes, src, begin, end, count, origin, inv, bvh.wnodes);
    if(launches) (*launches)++;
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, count, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    begin = end;
    end   = total;
    levels++;
    if(end > n || levels > 4096)
    {
      err = "collapseBvhWide: node count out of range";
      goto fail;
    }
  }
  bvh.numWideNodes = end;
  bvh.wideDepth    = (int)levels;
  tmpFree(src, stream); tmpFree(count, stream);
  return true;
fail:
  tmpFree(src, stream); tmpFree(count, stream); cudaFree(bvh.wnodes);
  bvh.wnodes = nullptr;
  return false;
}

// PLOC driver: fills children / ranges / parentInner / parentLeaf (hierarchyKernel conventions) and sortedOut (leaf order).
static bool plocTopology(cudaStream_t stream, const float4* wtri, const uint32_t* sortedVals, int n, int2* children, int2* ranges,
                         int* parentInner, int* parentLeaf, uint32_t* sortedOut, std::string& err, uint64_t* launches, int* roundsOut)
{
  int *cid = nullptr, *cidT = nullptr, *csize = nullptr, *csizeT = nullptr, *nn = nullptr, *nodeSize = nullptr, *counters = nullptr;
  int *leafPos = nullptr, *parentLeafOld = nullptr;
  Box *cbox = nullptr, *cboxT = nullptr;
  uint32_t *valid = nullptr, *pos = nullptr;
  void*     scanTemp = nullptr;
  size_t    scanBytes = 0;
  const int B = 256;
  int       c = n, rounds = 0, forcePair = 0;
  CK(tmpAlloc(&cid, 4ull * n, stream)); CK(tmpAlloc(&cidT, 4ull * n, stream)); CK(tmpAlloc(&csize, 4ull * n, stream)); CK(tmpAlloc(&csizeT, 4ull * n, stream));
  CK(tmpAlloc(&nn, 4ull * n, stream)); CK(tmpAlloc(&nodeSize, 4ull * n, stream)); CK(tmpAlloc(&counters, 8, stream));
  CK(tmpAlloc(&leafPos, 4ull * n, stream)); CK(tmpAlloc(&parentLeafOld, 4ull * n, stream));
  CK(tmpAlloc(&cbox, sizeof(Box) * (size_t)n, stream)); CK(tmpAlloc(&cboxT, sizeof(Box) * (size_t)n, stream));
  CK(tmpAlloc(&valid, 4ull * n, stream)); CK(tmpAlloc(&pos, 4ull * n, stream));
  CK(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, valid, pos, n, stream));
  CK(tmpAlloc(&scanTemp, scanBytes ? scanBytes : 16, stream));
  CK(cudaMemsetAsync(counters, 0, 8, stream));
  plocInitKernel<<<(n + B - 1) / B, B, 0, stream>>>(wtri, sortedVals, n, cid, cbox, csize);
  if(launches) (*launches)++;
  while(c > 1)
  {
    const int grid = (c + B - 1) / B;
    plocNearestKernel<<<grid, B, 0, stream>>>(cbox, c, nn, forcePair, plocRadius(c));
    plocMergeKernel<<<grid, B, 0, stream>>>(cid, cbox, csize, nn, c, n, counters, children, parentInner, parentLeafOld, nodeSize, cidT, cboxT,
                                             csizeT, valid);
    CK(cub::DeviceScan::ExclusiveSum(scanTemp, scanBytes, valid, pos, c, stream));
    plocCompactKernel<<<grid, B, 0, stream>>>(valid, pos, c, cidT, cboxT, csizeT, cid, cbox, csize, counters + 1);
    if(launches) (*launches) += 4;
    int next = 0;
    CK(cudaMemcpyAsync(&next, counters + 1, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if(next < 1 || next >= c || ++rounds > 100000)
    {
      err = "plocTopology: no progress";
      goto fail;
    }
    // a round normally removes 30-40 % of the clusters.  Only a degenerate input (thousands of identical boxes) merges one pair
    // per round; with many clusters left that would take forever, so neighbours are paired up for one round instead
    forcePair = (c > 1024 && (long long)(c - next) * 256 < c) ? 1 : 0;
    c         = next;
  }
  {
    const int root = -1;
    CK(cudaMemcpyAsync(parentInner, &root, 4, cudaMemcpyHostToDevice, stream));  // node 0 is the root
    const int grid = (n + B - 1) / B;
    plocOrderKernel<<<grid, B, 0, stream>>>(n, children, parentInner, parentLeafOld, nodeSize, leafPos, ranges);
    plocRemapKernel<<<grid, B, 0, stream>>>(n, children, leafPos, sortedVals, sortedOut, parentLeafOld, parentLeaf);
    if(launches) (*launches) += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
  }
  if(roundsOut) *roundsOut = rounds;
  tmpFree(cid, stream); tmpFree(cidT, stream); tmpFree(csize, stream); tmpFree(csizeT, stream); tmpFree(nn, stream); tmpFree(nodeSize, stream); tmpFree(counters, stream);
  tmpFree(leafPos, stream); tmpFree(parentLeafOld, stream); tmpFree(cbox, stream); tmpFree(cboxT, stream); tmpFree(valid, stream); tmpFree(pos, stream); tmpFree(scanTemp, stream);
  return true;
fail:
  tmpFree(cid, stream); tmpFree(cidT, stream); tmpFree(csize, stream); tmpFree(csizeT, stream); tmpFree(nn, stream); tmpFree(nodeSize, stream); tmpFree(counters, stream);
  tmpFree(leafPos, stream); tmpFree(parentLeafOld, stream); tmpFree(cbox, stream); tmpFree(cboxT, stream); tmpFree(valid, stream); tmpFree(pos, stream); tmpFree(scanTemp, stream);
  return false;
}

// Path-code keys of the GPU binned-SAH build (see 2b): fills keys / vals like mortonKernel.  No host synchronisation: every
// level is launched unconditionally (kSahPathBits + 1 of them; the node counts live on the device, levels past the real depth of the
// tree find nothing to do) and the temporaries come from the stream-ordered pool.
static bool sahPathKeys(cudaStream_t stream, const float4* wtri, uint32_t n, const uint32_t* boundsOrdered, uint64_t* keys, uint32_t* vals,
                        std::string& err, uint64_t* launches)
{
  // nodes of <= stop triangles are closed without a look at the bins (measured on the 263 k-triangle scene, closest + shadow ms per 32
  // frames: stop 2: 19.5 + 14.0 = the host builder's tree; 4: 19.9 + 14.3; 8: 20.3 + 14.5).  Huge scenes raise it so that the bins
  // (2688 bytes per open node) stay below 2 GB.
  static const uint32_t stopEnv = []() { const char* e = getenv("VKRT_SAH_STOP"); int v = e ? atoi(e) : 2; return (uint32_t)(v < 1 ? 1 : (v > 64 ? 64 : v)); }();
  uint32_t stop = stopEnv;
  while((size_t)(n / (stop + 1u) + 1u) * 3 * kSahBins * sizeof(SahBin) > ((size_t)2 << 30) && stop < 64u)
    stop++;
  const uint32_t maxOpen  = n / (stop + 1u) + 1u;  // open nodes hold > stop triangles each
  const uint32_t maxSlots = 2u * maxOpen + 2u;     // ... and the nodes of a level are the children of the previous level's open nodes
  SahSlot *slotsA = nullptr, *slotsB = nullptr;
  SahBin*   bins = nullptr;
  SahLevel* levels = nullptr;
  int32_t*  triSlot = nullptr;
  uint32_t *triCode = nullptr, *triDepth = nullptr, *triFinal = nullptr, *finalCount = nullptr;
  Box*      finals = nullptr;
  const int B = 256;
  const uint32_t grid = (n + B - 1) / B;
  TCK(tmpAlloc(&slotsA, sizeof(SahSlot) * (size_t)maxSlots, stream)); TCK(tmpAlloc(&slotsB, sizeof(SahSlot) * (size_t)maxSlots, stream));
  TCK(tmpAlloc(&bins, sizeof(SahBin) * (size_t)maxOpen * 3 * kSahBins, stream));
  TCK(tmpAlloc(&levels, sizeof(SahLevel) * (kSahPathBits + 2), stream));
  TCK(tmpAlloc(&triSlot, 4ull * n, stream)); TCK(tmpAlloc(&triCode, 4ull * n, stream)); TCK(tmpAlloc(&triDepth, 4ull * n, stream));
  TCK(tmpAlloc(&triFinal, 4ull * n, stream));
  TCK(tmpAlloc(&finalCount, 4, stream)); TCK(tmpAlloc(&finals, sizeof(Box) * (size_t)(n + 1), stream));
  TCK(cudaMemsetAsync(triSlot, 0, 4ull * n, stream)); TCK(cudaMemsetAsync(triCode, 0, 4ull * n, stream));
  TCK(cudaMemsetAsync(triDepth, 0, 4ull * n, stream)); TCK(cudaMemsetAsync(triFinal, 0, 4ull * n, stream));
  sahRootKernel<<<1, 1, 0, stream>>>(slotsA, boundsOrdered, n, levels, finalCount, stop);
  if(launches) (*launches)++;
  for(int level = 0; level <= kSahPathBits; level++)
  {
    // upper bounds for the grids: 2^level nodes at most, never more than maxSlots
    const uint32_t slotsBound = (level < 31 && (1u << level) < maxSlots) ? (1u << level) : maxSlots;
    const uint32_t openBound  = std::min<uint32_t>(slotsBound, maxOpen);
    const size_t   binBound   = (size_t)openBound * 3 * kSahBins;
    sahBinClearKernel<<<(unsigned)std::min<size_t>((binBound + B - 1) / B, 148 * 16), B, 0, stream>>>(bins, levels + level);
    sahBinKernel<<<grid, B, 0, stream>>>(wtri, n, triSlot, slotsA, bins);
    sahSplitKernel<<<(unsigned)(((size_t)slotsBound * 32 + B - 1) / B), B, 0, stream>>>(slotsA, levels, level, bins, slotsB, finals, finalCount, stop, maxSlots);
    sahApplyKernel<<<grid, B, 0, stream>>>(wtri, n, triSlot, triCode, triDepth, triFinal, slotsA, slotsB);
    if(launches) (*launches) += 4;
    std::swap(slotsA, slotsB);
  }
  sahKeyKernel<<<grid, B, 0, stream>>>(wtri, n, triCode, triDepth, triFinal, finals, keys, vals);
  if(launches) (*launches)++;
  TCK(cudaGetLastError());
  tmpFree(slotsA, stream); tmpFree(slotsB, stream); tmpFree(bins, stream); tmpFree(levels, stream); tmpFree(triSlot, stream); tmpFree(triCode, stream);
  tmpFree(triDepth, stream); tmpFree(triFinal, stream); tmpFree(finalCount, stream); tmpFree(finals, stream);
  return true;
fail:
  tmpFree(slotsA, stream); tmpFree(slotsB, stream); tmpFree(bins, stream); tmpFree(levels, stream); tmpFree(triSlot, stream); tmpFree(triCode, stream);
  tmpFree(triDepth, stream); tmpFree(triFinal, stream); tmpFree(finalCount, stream); tmpFree(finals, stream);
  return false;
}

// Host driver.  All temporaries are freed before returning; nodes/tris/triInfo are owned by the caller.
bool buildBvh(cudaStream_t stream, const InstanceXform* dInst, const uint32_t* dInstTriOffset, uint32_t numInst, const InstanceData* dGeo,
              uint32_t numTris, int buildMode, BvhResult& out, std::string& err, uint64_t* launches)
{
  out = BvhResult{};
  float4*   wtri = nullptr;
  uint32_t* bounds = nullptr;
  uint64_t *keysA = nullptr, *keysB = nullptr;
  uint32_t *valsA = nullptr, *valsB = nullptr;
  void*     cubTemp = nullptr;
  int2 *children = nullptr, *ranges = nullptr;
  int *parentInner = nullptr, *parentLeaf = nullptr, *flags = nullptr, *nodeDepth = nullptr, *maxDepthDev = nullptr;
  Box *leafBox = nullptr, *nodeBox = nullptr;
  float*    nodeCost = nullptr;
  uint32_t* collapsed = nullptr;
  const int B = 256;
  uint32_t  numNodes = numTris > 1 ? numTris - 1 : 1;

  // nodes and triangles share one allocation so that a single L2 access-policy window can cover the BVH
  CK(cudaMalloc(&out.nodes, sizeof(float4) * (4ull * numNodes + 3ull * (numTris ? numTris : 1))));
  out.tris       = out.nodes + 4ull * numNodes;
  out.arenaBytes = sizeof(float4) * (4ull * numNodes + 3ull * (numTris ? numTris : 1));
  CK(cudaMalloc(&out.triInfo, sizeof(uint2) * (numTris ? numTris : 1)));
  CK(cudaMalloc(&out.triShade, sizeof(uint4) * 7ull * (numTris ? numTris : 1)));
  out.numNodes = numNodes;
  out.numTris  = numTris;
  if(numTris == 0)
  {
    emptySceneKernel<<<1, 1, 0, stream>>>(out.nodes);
    if(launches) (*launches)++;
    CK(cudaStreamSynchronize(stream));
    return true;
  }
  {
    size_t n = numTris;
    CK(tmpAlloc(&wtri, sizeof(float4) * 3 * n, stream));
    CK(tmpAlloc(&bounds, 6 * sizeof(uint32_t), stream));
    CK(tmpAlloc(&keysA, 8 * n, stream)); CK(tmpAlloc(&keysB, 8 * n, stream));
    CK(tmpAlloc(&valsA, 4 * n, stream)); CK(tmpAlloc(&valsB, 4 * n, stream));
    CK(tmpAlloc(&children, sizeof(int2) * n, stream)); CK(tmpAlloc(&ranges, sizeof(int2) * n, stream));
    CK(tmpAlloc(&parentInner, 4 * n, stream)); CK(tmpAlloc(&parentLeaf, 4 * n, stream)); CK(tmpAlloc(&flags, 4 * n, stream));
    CK(tmpAlloc(&nodeDepth, 4 * n, stream)); CK(tmpAlloc(&maxDepthDev, 4, stream));
    CK(tmpAlloc(&leafBox, sizeof(Box) * n, stream)); CK(tmpAlloc(&nodeBox, sizeof(Box) * n, stream));
    CK(tmpAlloc(&nodeCost, 4 * n, stream)); CK(tmpAlloc(&collapsed, 4 * n, stream));

    uint32_t initB[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CK(cudaMemcpyAsync(bounds, initB, sizeof(initB), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(flags, 0, 4 * n, stream));
    CK(cudaMemsetAsync(collapsed, 0, 4 * n, stream));
    CK(cudaMemsetAsync(maxDepthDev, 0, 4, stream));
    uint32_t grid = (uint32_t)((n + B - 1) / B);
    flattenKernel<<<grid, B, 0, stream>>>(dInst, dInstTriOffset, numInst, dGeo, numTris, wtri, out.triInfo, out.triShade, bounds);
    if(buildMode == VKRT_BUILD_SAH_GPU && n > 1)
    {
      if(!sahPathKeys(stream, wtri, numTris, bounds, keysA, valsA, err, launches))
        goto fail;
    }
    else
      mortonKernel<<<grid, B, 0, stream>>>(wtri, numTris, bounds, keysA, valsA);
    size_t tempBytes = 0;
    cub::DoubleBuffer<uint64_t> dk(keysA, keysB);
    cub::DoubleBuffer<uint32_t> dv(valsA, valsB);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, dk, dv, (int)n, 0, 63, stream));
    CK(tmpAlloc(&cubTemp, tempBytes ? tempBytes : 16, stream));
    CK(cub::DeviceRadixSort::SortPairs(cubTemp, tempBytes, dk, dv, (int)n, 0, 63, stream));
    const uint64_t* sk = dk.Current();
    const uint32_t* sv = dv.Current();
    if(n > 1 && buildMode == VKRT_BUILD_PLOC)
    {
      // the other key buffer is free after the sort: it receives the depth-first leaf order
      uint32_t* svOut = (dv.Current() == valsA) ? valsB : valsA;
      if(!plocTopology(stream, wtri, sv, (int)n, children, ranges, parentInner, parentLeaf, svOut, err, launches, nullptr))
        goto fail;
      sv = svOut;
    }
    else if(n > 1)
      hierarchyKernel<<<grid, B, 0, stream>>>(sk, (int)n, children, ranges, parentInner, parentLeaf);
    int collapseMax = (buildMode == VKRT_BUILD_LBVH_COLLAPSED || buildMode == VKRT_BUILD_PLOC) ? 8 : 1;
    if(buildMode == VKRT_BUILD_SAH_GPU)
    {
      // leaves of the SAH tree hold up to 4 triangles like the host builder's (7.8 triangle tests per ray with 8, 5.9 with 4)
      static const int leafMax = []() { const char* e = getenv("VKRT_SAH_LEAF"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
      collapseMax = leafMax;
    }
    refitKernel<<<grid, B, 0, stream>>>(wtri, sv, (int)n, out.tris, children, parentInner, parentLeaf, leafBox, nodeBox, nodeCost,
                                       nodeDepth, collapsed, ranges, flags, collapseMax, maxDepthDev,
                                       (buildMode == VKRT_BUILD_SAH_GPU && n > 1) ? sk : nullptr);
    emitKernel<<<grid, B, 0, stream>>>((int)n, children, ranges, leafBox, nodeBox, collapsed, out.nodes);
    if(launches) (*launches) += 5 + 7;  // ours + the radix-sort passes of the library
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&out.maxDepth, maxDepthDev, 4, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }
  tmpFree(wtri, stream); tmpFree(bounds, stream); tmpFree(keysA, stream); tmpFree(keysB, stream); tmpFree(valsA, stream); tmpFree(valsB, stream); tmpFree(cubTemp, stream);
  tmpFree(children, stream); tmpFree(ranges, stream); tmpFree(parentInner, stream); tmpFree(parentLeaf, stream); tmpFree(flags, stream); tmpFree(nodeDepth, stream);
  tmpFree(maxDepthDev, stream); tmpFree(leafBox, stream); tmpFree(nodeBox, stream); tmpFree(nodeCost, stream); tmpFree(collapsed, stream);
  return true;
fail:
  tmpFree(wtri, stream); tmpFree(bounds, stream); tmpFree(keysA, stream); tmpFree(keysB, stream); tmpFree(valsA, stream); tmpFree(valsB, stream); tmpFree(cubTemp, stream);
  tmpFree(children, stream); tmpFree(ranges, stream); tmpFree(parentInner, stream); tmpFree(parentLeaf, stream); tmpFree(flags, stream); tmpFree(nodeDepth, stream);
  tmpFree(maxDepthDev, stream); tmpFree(leafBox, stream); tmpFree(nodeBox, stream); tmpFree(nodeCost, stream); tmpFree(collapsed, stream);
  cudaFree(out.nodes); cudaFree(out.triInfo); cudaFree(out.triShade);
  out = BvhResult{};
  return false;
}

}  // namespace vkrt
