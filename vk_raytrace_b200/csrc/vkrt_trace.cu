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
//   * nodes are 64 B fetched as 2 x ld.global.nc.v8 (centre / half-extent boxes: the slab test is FMA-pipe work),
//     triangles 48 B as 3 x ld.global.nc.v4; optional formats: min/max fp32 nodes, 32-byte 16-bit grid nodes, and a
//     4-wide tree with 64-byte quantised nodes (traceKernelWide, the default above 1.5 M triangles).
//
// Result semantics (DESIGN.md "traversal semantics", mirrored by oracle/o_scene.h):
// nearest Moller-Trumbore hit with t in (0, tmax), ties to the lower triangle id,
// back faces culled unless the instance is double sided, non-opaque candidates pass
// the hashed alpha test.  Box culling is conservative, so the result does not depend
// on tree shape, visiting order or scheduling.
#include "vkrt_device.cuh"
#include "vkrt_texture.cuh"
#include <cstdlib>
#include <cstring>
#include "vkrt_internal.h"

namespace vkrt {

#ifndef VKRT_STACK_SHARED
#define VKRT_STACK_SHARED 32
#endif
#ifndef VKRT_REFETCH
#define VKRT_REFETCH 24
#endif
#ifndef VKRT_TRACE_MINBLOCKS
#define VKRT_TRACE_MINBLOCKS 8
#endif
#ifndef VKRT_WIDE_MINBLOCKS
#define VKRT_WIDE_MINBLOCKS 8
#endif
#ifndef VKRT_TRACE_BLOCK
#define VKRT_TRACE_BLOCK 128
#endif
#ifndef VKRT_LEAF_STREAM
#define VKRT_LEAF_STREAM 1  // 0: one parked leaf per pass of the leaf loop (round 1)
#endif
#ifndef VKRT_BRANCHFREE
#define VKRT_BRANCHFREE 0  // 1: branch-free push / pop / park (selects over the two topmost stack entries).  Measured on B200 (gpurun_out/bench_r02a_*):
                           // 15 % fewer issue slots but a longer dependent chain per step -> closest-hit 19.96 vs 19.05 ms per 32 frames: off
#endif
constexpr int      kTraceBlock       = VKRT_TRACE_BLOCK;
constexpr int      kStackShared      = VKRT_STACK_SHARED;  // entries per thread in shared memory
constexpr int      kStackLocal       = 128 - VKRT_STACK_SHARED;  // overflow entries in local memory
constexpr int      kRefetchThreshold = VKRT_REFETCH;  // refill the warp when fewer lanes than this are traversing
constexpr uint32_t kDone             = 0x7fffffffu;  // traversal finished (never a node index)
constexpr unsigned kFull             = 0xffffffffu;

enum TraceMode { MODE_CLOSEST = 0, MODE_SHADOW_RESOLVE = 1, MODE_ANY_PROBE = 2 };

// ---- alpha test for non-FORCE_OPAQUE instances (cold path) --------------------------------
__device__ __noinline__ bool alphaTestPasses(const SceneView& sc, uint32_t triId, uint32_t matId, float bu, float bv, uint32_t raySeed)
{
  // two independent gathers (the material's alpha record, the triangle's texture coordinates), then the texels
  const AlphaMat& am = sc.alphaMats[matId];
  const uint4*    ts = sc.triShade + 7ull * triId;  // {pos, n | uv, tangent, colour} x 3, see flattenKernel
  const uint4     a1 = __ldg(ts + 1), b1 = __ldg(ts + 3), c1 = __ldg(ts + 5);
  float           baseColorAlpha = am.alphaFactor;
  if(am.hasTexture)
  {
    // texcoord as stored (the tangent handedness bit in v's LSB included, as traceray_rq.glsl:72-83 reads it)
    const float b0 = 1.0f - bu - bv;
    float2      tc = f2(__uint_as_float(a1.x), __uint_as_float(a1.y)) * b0 + f2(__uint_as_float(b1.x), __uint_as_float(b1.y)) * bu
                + f2(__uint_as_float(c1.x), __uint_as_float(c1.y)) * bv;
    tc             = applyUvTransform(am.uvRows, tc);
    baseColorAlpha *= sampleTexture(am.texture, tc).w;
  }
  float opacity = (am.alphaMode == VKRT_ALPHA_MASK) ? (baseColorAlpha > am.alphaCutoff ? 1.0f : 0.0f) : baseColorAlpha;
  return !(alphaDraw(raySeed, triId) > opacity);
}

// One 64-byte node as two 256-bit loads (sm_100 LDG.E.256) instead of four 128-bit ones: fewer load instructions on the
// step's dependency chain (the kernel is bound by the latency of these fetches, profiles/README.md).
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

// One 32-byte quantised node (quantizeBvh) as a single 256-bit load.
VKRT_D void loadQNode(const uint4* __restrict__ qnodes, uint32_t node, uint32_t (&w)[8])
{
  const uint4* p = qnodes + 2ull * node;
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
      : "l"(p));
}
// 2^23 + q as a float: the 16-bit grid coordinate dropped into the mantissa of 0x4B000000 by one byte permute;
// sel = 0x7610 picks the low half-word (box min), 0x7632 the high one (box max)
VKRT_D float gridPlane(uint32_t w, uint32_t sel) { return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)); }

VKRT_D float safeInv(float x) { return 1.0f / (fabsf(x) < 1e-20f ? copysignf(1e-20f, x) : x); }

__device__ __noinline__ void stackSpill(uint32_t* l, int i, uint32_t v) { l[min(i, kStackLocal - 1)] = v; }
__device__ __noinline__ uint32_t stackFill(const uint32_t* l, int i) { return l[min(i, kStackLocal - 1)]; }

enum NodeFormat { FMT_LOHI = 0, FMT_Q32 = 1, FMT_CE = 2 };

template <int MODE, bool COUNT, int VARIANT, int FMT>
__global__ void __launch_bounds__(kTraceBlock, VKRT_TRACE_MINBLOCKS) traceKernel(const __grid_constant__ SceneView sc, const __grid_constant__ TraceLaunch L)
{
  __shared__ uint32_t s_stack[(kStackShared > 0 ? kStackShared : 1) * (kStackShared > 0 ? kTraceBlock : 1)];
  uint32_t* const sstack = &s_stack[threadIdx.x];
  uint32_t        lstack[kStackLocal];
  const uint32_t  n    = *L.count;
  if(MODE == MODE_SHADOW_RESOLVE && L.stats && blockIdx.x == 0 && threadIdx.x == 0)
  {
    atomicAdd(&L.stats->shadowRays, (unsigned long long)n);  // AnyHit() calls of this bounce (round 1 counted them in a kernel of their own)
    if(L.primary)
      atomicAdd(&L.stats->shadowRaysPrimary, (unsigned long long)n);
  }
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

  const float4* __restrict__ nodes  = sc.nodes;
  const uint4* __restrict__  qnodes = sc.qnodes;
  const float4* __restrict__ cnodes = sc.cnodes;
  constexpr bool QN = (FMT == FMT_Q32);
  constexpr bool BRANCHFREE = (kStackShared >= 4) && (VKRT_BRANCHFREE != 0);
  const float4* __restrict__ tris   = sc.tris;

  // per-lane ray + traversal state (lives across refills)
  bool     haveRay = false;
  uint32_t slot = 0;
  float3   o = f3(0.f), d = f3(0.f), id = f3(0.f);
  // QN: id is the inverse direction in grid units, nood = t0 - (o_g + 2^23) * id, selN/selF pick the near/far plane
  float3   nood = f3(0.f);
  uint32_t selNx = 0, selNy = 0, selNz = 0;
  float    bestT = 0.f, bestU = 0.f, bestV = 0.f;
  uint32_t bestTri = ~0u;
  uint32_t node = kDone, leaf = 0;
  int      sp = 0;
  bool     exhausted = false;  // warp uniform: the queue has been handed out completely
  uint32_t nNodes = 0, nTris = 0, rayBase = 0;

  // the overflow half of the stack sits behind real calls, so that the common path is one predicated shared-memory access
  auto push = [&](uint32_t v) {
    if(sp < kStackShared)
      sstack[sp * kTraceBlock] = v;
    else
      stackSpill(lstack, sp - kStackShared, v);
    sp++;
  };
  auto pop = [&]() -> uint32_t {
    if(sp == 0)
      return kDone;
    --sp;
    if(sp < kStackShared)
      return sstack[sp * kTraceBlock];
    return stackFill(lstack, sp - kStackShared);
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
          slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? ldS(&L.queue[i]) : i;
          const float4 ro = ldS(&L.rayO[slot]), rd = ldS(&L.rayD[slot]);
          o = f3(ro); d = f3(rd);
          const float tmax = ro.w;
          bestT   = tmax; bestU = 0.f; bestV = 0.f; bestTri = ~0u;
          node    = 0; leaf = 0; sp = 0;
          haveRay = true;
          rayBase = (COUNT && L.cost) ? deviceNs() : 0u;  // heat map: when this lane took the ray
          if(QN)
          {
            // the ray in grid space; its origin is moved to (just before) the entry into the grid cube so that
            // o_g + 2^23 keeps whole-cell resolution however far away the ray starts.  Box times stay world times.
            const float3 dg = d * sc.gridInv;
            id              = f3(safeInv(dg.x), safeInv(dg.y), safeInv(dg.z));
            float3      og  = (o - sc.gridOrigin) * sc.gridInv;
            const float ax = (0.0f - og.x) * id.x, bx = (65535.0f - og.x) * id.x;
            const float ay = (0.0f - og.y) * id.y, by = (65535.0f - og.y) * id.y;
            const float az = (0.0f - og.z) * id.z, bz = (65535.0f - og.z) * id.z;
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
            if(!(tn <= tf * 1.0002f))
              node = kDone;  // misses the whole scene
            const float t0 = tn * 0.99995f;
            og             = og + dg * t0;
            nood  = f3(t0 - (og.x + 8388608.0f) * id.x, t0 - (og.y + 8388608.0f) * id.y, t0 - (og.z + 8388608.0f) * id.z);
            selNx = (__float_as_uint(id.x) >> 31) ? 0x7632u : 0x7610u;
            selNy = (__float_as_uint(id.y) >> 31) ? 0x7632u : 0x7610u;
            selNz = (__float_as_uint(id.z) >> 31) ? 0x7632u : 0x7610u;
          }
          else
          {
            id = f3(safeInv(d.x), safeInv(d.y), safeInv(d.z));
            if(FMT == FMT_CE)
              nood = f3(-(o.x * id.x), -(o.y * id.y), -(o.z * id.z));
          }
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
        // s1 / s2: the two topmost stack entries as of the start of an iteration (branch-free step; junk below the bottom,
        // guarded by sp where they are used).  The leaf phase and the refill change the stack, so they are re-read here.
        uint32_t s1 = 0, s2 = 0;
        if(BRANCHFREE)
        {
          s1 = sstack[max(min(sp, kStackShared) - 1, 0) * kTraceBlock];
          s2 = sstack[max(min(sp, kStackShared) - 2, 0) * kTraceBlock];
        }
        while(node < kDone)  // inner index (leaf codes have bit 31 set, kDone is larger than any index)
        {
          float    tn0, tf0, tn1, tf1;
          uint32_t c0, c1;
          if(COUNT) nNodes++;
          if(QN)
          {
            uint32_t w[8];
            loadQNode(qnodes, node, w);
            // far selectors are re-derived per node (volatile: keeps them out of registers that would otherwise spill)
            uint32_t selFx, selFy, selFz;
            asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(selFx) : "r"(selNx));
            asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(selFy) : "r"(selNy));
            asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(selFz) : "r"(selNz));
            tn0 = fmaxf(fmaxf(__fmaf_rn(gridPlane(w[0], selNx), id.x, nood.x), __fmaf_rn(gridPlane(w[1], selNy), id.y, nood.y)),
                        fmaxf(__fmaf_rn(gridPlane(w[2], selNz), id.z, nood.z), 0.0f));
            tf0 = fminf(fminf(__fmaf_rn(gridPlane(w[0], selFx), id.x, nood.x), __fmaf_rn(gridPlane(w[1], selFy), id.y, nood.y)),
                        fminf(__fmaf_rn(gridPlane(w[2], selFz), id.z, nood.z), bestT));
            tn1 = fmaxf(fmaxf(__fmaf_rn(gridPlane(w[3], selNx), id.x, nood.x), __fmaf_rn(gridPlane(w[4], selNy), id.y, nood.y)),
                        fmaxf(__fmaf_rn(gridPlane(w[5], selNz), id.z, nood.z), 0.0f));
            tf1 = fminf(fminf(__fmaf_rn(gridPlane(w[3], selFx), id.x, nood.x), __fmaf_rn(gridPlane(w[4], selFy), id.y, nood.y)),
                        fminf(__fmaf_rn(gridPlane(w[5], selFz), id.z, nood.z), bestT));
            c0 = w[6];
            c1 = w[7];
          }
          else if(FMT == FMT_CE)
          {
            // centre / half-extent boxes (centerNodesKernel): everything up to the 3-input min/max runs on the FMA pipe
            float4 m0, m1, m2, m3;
            loadNode256(cnodes, node, m0, m1, m2, m3);
            const float c0x = __fmaf_rn(m0.x, id.x, nood.x), c0y = __fmaf_rn(m0.z, id.y, nood.y), c0z = __fmaf_rn(m2.x, id.z, nood.z);
            const float c1x = __fmaf_rn(m1.x, id.x, nood.x), c1y = __fmaf_rn(m1.z, id.y, nood.y), c1z = __fmaf_rn(m2.z, id.z, nood.z);
            const float aix = fabsf(id.x), aiy = fabsf(id.y), aiz = fabsf(id.z);
            tn0 = fmaxf(fmaxf(__fmaf_rn(-m0.y, aix, c0x), __fmaf_rn(-m0.w, aiy, c0y)), __fmaf_rn(-m2.y, aiz, c0z));
            tf0 = fminf(fminf(__fmaf_rn(m0.y, aix, c0x), __fmaf_rn(m0.w, aiy, c0y)), __fmaf_rn(m2.y, aiz, c0z));
            tn1 = fmaxf(fmaxf(__fmaf_rn(-m1.y, aix, c1x), __fmaf_rn(-m1.w, aiy, c1y)), __fmaf_rn(-m2.w, aiz, c1z));
            tf1 = fminf(fminf(__fmaf_rn(m1.y, aix, c1x), __fmaf_rn(m1.w, aiy, c1y)), __fmaf_rn(m2.w, aiz, c1z));
            tf0 = fminf(tf0, bestT);
            tf1 = fminf(tf1, bestT);
            c0  = __float_as_uint(m3.x);
            c1  = __float_as_uint(m3.y);
          }
          else
          {
          float4 n0, n1, n2, n3;
          loadNode256(nodes, node, n0, n1, n2, n3);
          const float a0x = (n0.x - o.x) * id.x, b0x = (n0.y - o.x) * id.x;
          const float a0y = (n0.z - o.y) * id.y, b0y = (n0.w - o.y) * id.y;
          const float a0z = (n2.x - o.z) * id.z, b0z = (n2.y - o.z) * id.z;
          const float a1x = (n1.x - o.x) * id.x, b1x = (n1.y - o.x) * id.x;
          const float a1y = (n1.z - o.y) * id.y, b1y = (n1.w - o.y) * id.y;
          const float a1z = (n2.z - o.z) * id.z, b1z = (n2.w - o.z) * id.z;
          tn0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
          tf0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), bestT));
          tn1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
          tf1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), bestT));
          c0  = __float_as_uint(n3.x);
          c1  = __float_as_uint(n3.y);
          }
          // conservative: never reject a box the triangle test could hit.  FMT_CE leaves tn unclamped: max(tn,0) <= x is tn <= x && 0 <= x
          const float x0 = tf0 * 1.0002f, x1 = tf1 * 1.0002f;
          const bool  h0 = (FMT == FMT_CE) ? (tn0 <= x0 && x0 >= 0.0f) : (tn0 <= x0);
          const bool  h1 = (FMT == FMT_CE) ? (tn1 <= x1 && x1 >= 0.0f) : (tn1 <= x1);
          if(BRANCHFREE && sp < kStackShared)
          {
            // Branch-free step (round 2; ncu of round 1: the push / pop / park branches below ran with 2-4 lanes each and cost as
            // many issue slots as the 21-lane box test).  The two topmost stack entries were fetched at the top of the iteration
            // (the address depends on sp only, so the loads are off the node-fetch dependency chain); push, pop and leaf
            // parking become selects over {far child, top, second}:
            //   push  far child when both boxes are hit (the store always happens: slot sp is free)
            //   pop   when neither is hit
            //   park  when the next node is a leaf and none is parked yet: take the following stack entry
            const bool     both = h0 && h1, none = !(h0 || h1);
            const bool     sw   = both && (tn1 < tn0);
            const uint32_t nearC = (h0 && !sw) ? c0 : c1;
            const uint32_t farC  = sw ? c0 : c1;
            sstack[sp * kTraceBlock] = farC;
            const uint32_t t1 = both ? farC : s1, t2 = both ? s1 : s2;
            int            spn = sp + (both ? 1 : 0);
            const bool     popped = none && spn > 0;
            const uint32_t nd = none ? (popped ? t1 : kDone) : nearC;
            const uint32_t u1 = popped ? t2 : t1;
            spn -= popped ? 1 : 0;
            const bool park = (nd & kLeafBit) && leaf == 0;
            leaf            = park ? nd : leaf;
            node            = park ? (spn > 0 ? u1 : kDone) : nd;
            sp              = spn - ((park && spn > 0) ? 1 : 0);
            // the next iteration's view of the stack top, fetched now so that the shared-memory latency overlaps the node fetch
            s1 = sstack[max(min(sp, kStackShared) - 1, 0) * kTraceBlock];
            s2 = sstack[max(min(sp, kStackShared) - 2, 0) * kTraceBlock];
          }
          else
          {
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
          if(BRANCHFREE)
          {
            s1 = sstack[max(min(sp, kStackShared) - 1, 0) * kTraceBlock];
            s2 = sstack[max(min(sp, kStackShared) - 2, 0) * kTraceBlock];
          }
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
#if VKRT_LEAF_STREAM
        // One loop over the triangles of the parked leaf and, right behind them, of a second leaf the lane found while it kept
        // descending (node then holds that leaf's code): the warp iterates max(cnt1 + cnt2) times instead of max(cnt1) + max(cnt2)
        // (ncu, round 1: 7.7 of 32 lanes active in this loop for incoherent rays -- leaves hold 1..8 triangles).
        if(leaf != 0)
        {
          uint32_t cur = leaf & kLeafMask, end = cur + ((leaf >> kLeafCntSh) & 7u) + 1u;
          for(;;)
          {
            if(cur == end)
            {
              if(!(node & kLeafBit))
                break;
              leaf = node;
              node = pop();
              cur  = leaf & kLeafMask;
              end  = cur + ((leaf >> kLeafCntSh) & 7u) + 1u;
            }
            const float4 q0 = __ldg(&tris[3ull * cur + 0]);
            const float4 q1 = __ldg(&tris[3ull * cur + 1]);
            const float4 q2 = __ldg(&tris[3ull * cur + 2]);
            cur++;
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
            // t in (0, tmax): bestT starts at tmax, so only the tie branch has to exclude "no hit yet"
            if(!(t > 0.0f && (t < bestT || (t == bestT && bestTri != ~0u && triId < bestTri))))
              continue;
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, __float_as_uint(q2.w), u, v, __float_as_uint(L.rayD[slot].w)))
              continue;
            bestT = t; bestU = u; bestV = v; bestTri = triId;
            if(MODE != MODE_CLOSEST)
            {
              node = kDone;  // TerminateOnFirstHit
              sp   = 0;
              break;
            }
          }
          leaf = 0;
        }
#else
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
            // t in (0, tmax): bestT starts at tmax, so only the tie branch has to exclude "no hit yet"
            if(!(t > 0.0f && (t < bestT || (t == bestT && bestTri != ~0u && triId < bestTri))))
              continue;
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, __float_as_uint(q2.w), u, v, __float_as_uint(L.rayD[slot].w)))
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

#endif

        // ---- dynamic fetch: leave to refill when the warp has thinned out ------------------------
        if(!exhausted && __popc(__activemask()) < kRefetchThreshold)
          break;
      }

      if(node == kDone && leaf == 0)
      {
        if(COUNT && L.cost && MODE != MODE_ANY_PROBE)
        {
          // heat-map debug mode: the nanoseconds this lane held the ray (waiting for the other lanes of its warp included, as in the
          // reference's per-invocation clock), charged to its path (one closest and one shadow ray per path and launch)
          const uint32_t path = (MODE == MODE_SHADOW_RESOLVE) ? __float_as_uint(L.shC[slot].w) : slot;
          L.cost[path].w += (float)(deviceNs() - rayBase);
        }
        if(MODE == MODE_CLOSEST)
          stS(&L.hit[slot], (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(bestT, bestU, bestV, __uint_as_float(bestTri)));
        else if(MODE == MODE_ANY_PROBE)
          L.hit[slot] = (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(0.f, 0.f, 0.f, __uint_as_float(0u));
        else if(bestTri == ~0u)
        {
          // unoccluded: radiance += contribution (one shadow ray per path per bounce: no race)
          const float4   c    = ldS(&L.shC[slot]);
          const uint32_t path = __float_as_uint(c.w);
          float4         rad  = ldS(&L.rad[path]);
          rad.x += c.x; rad.y += c.y; rad.z += c.z;
          stS(&L.rad[path], rad);
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


// ---------------------------------------------------------------------------------------------
// BVH4 traversal (collapseBvhWide): same persistent while-while structure, same triangle test, same results.
// One step = two 256-bit loads + four grid-space slab tests; the nearest hit child is entered, the others are pushed.
constexpr uint32_t kEmptyRef = 0xffffffffu;

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock, VKRT_WIDE_MINBLOCKS) traceKernelWide(const __grid_constant__ SceneView sc, const __grid_constant__ TraceLaunch L)
{
  __shared__ uint32_t s_stack[(kStackShared > 0 ? kStackShared : 1) * (kStackShared > 0 ? kTraceBlock : 1)];
  uint32_t* const sstack = &s_stack[threadIdx.x];
  uint32_t        lstack[kStackLocal];
  const uint32_t  n    = *L.count;
  if(MODE == MODE_SHADOW_RESOLVE && L.stats && blockIdx.x == 0 && threadIdx.x == 0)
  {
    atomicAdd(&L.stats->shadowRays, (unsigned long long)n);  // AnyHit() calls of this bounce (round 1 counted them in a kernel of their own)
    if(L.primary)
      atomicAdd(&L.stats->shadowRaysPrimary, (unsigned long long)n);
  }
  const uint32_t  lane = threadIdx.x & 31u;

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

  const uint4* __restrict__  wnodes = sc.wnodes;
  const float4* __restrict__ tris   = sc.tris;

  bool     haveRay = false;
  uint32_t slot = 0;
  float3   o = f3(0.f), d = f3(0.f), id = f3(0.f), nood = f3(0.f);
  uint32_t selNx = 0, selNy = 0, selNz = 0;
  float    bestT = 0.f, bestU = 0.f, bestV = 0.f;
  uint32_t bestTri = ~0u;
  uint32_t node = kDone, leaf = 0;
  int      sp = 0;
  bool     exhausted = false;
  uint32_t nNodes = 0, nTris = 0, rayBase = 0;

  // the overflow half of the stack sits behind real calls, so that the common path is one predicated shared-memory access
  auto push = [&](uint32_t v) {
    if(sp < kStackShared)
      sstack[sp * kTraceBlock] = v;
    else
      stackSpill(lstack, sp - kStackShared, v);
    sp++;
  };
  auto pop = [&]() -> uint32_t {
    if(sp == 0)
      return kDone;
    --sp;
    if(sp < kStackShared)
      return sstack[sp * kTraceBlock];
    return stackFill(lstack, sp - kStackShared);
  };

  for(;;)
  {
    __syncwarp();
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
        const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
        if(i < n)
        {
          slot = (MODE != MODE_SHADOW_RESOLVE && L.queue) ? ldS(&L.queue[i]) : i;
          const float4 ro = ldS(&L.rayO[slot]), rd = ldS(&L.rayD[slot]);
          o = f3(ro); d = f3(rd);
          const float tmax = ro.w;
          bestT   = tmax; bestU = 0.f; bestV = 0.f; bestTri = ~0u;
          node    = 0; leaf = 0; sp = 0;
          haveRay = true;
          rayBase = (COUNT && L.cost) ? deviceNs() : 0u;  // heat map: when this lane took the ray
          // grid-space ray, see traceKernel<QN>
          const float3 dg = d * sc.gridInv;
          id              = f3(safeInv(dg.x), safeInv(dg.y), safeInv(dg.z));
          float3      og  = (o - sc.gridOrigin) * sc.gridInv;
          const float ax = (0.0f - og.x) * id.x, bx = (65535.0f - og.x) * id.x;
          const float ay = (0.0f - og.y) * id.y, by = (65535.0f - og.y) * id.y;
          const float az = (0.0f - og.z) * id.z, bz = (65535.0f - og.z) * id.z;
          const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
          const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
          if(!(tn <= tf * 1.0002f))
            node = kDone;
          const float t0 = tn * 0.99995f;
          og             = og + dg * t0;
          nood  = f3(t0 - (og.x + 8388608.0f) * id.x, t0 - (og.y + 8388608.0f) * id.y, t0 - (og.z + 8388608.0f) * id.z);
          selNx = (__float_as_uint(id.x) >> 31) ? 0x7632u : 0x7610u;
          selNy = (__float_as_uint(id.y) >> 31) ? 0x7632u : 0x7610u;
          selNz = (__float_as_uint(id.z) >> 31) ? 0x7632u : 0x7610u;
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
        while(node < kDone)
        {
          uint32_t w[16];
          {
            const uint4* p = wnodes + 4ull * node;
            asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                : "l"(p));
            asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                : "l"(p + 2));
          }
          if(COUNT) nNodes++;
          // far selectors re-derived per step (volatile: three registers that would otherwise be spilled and reloaded here)
          uint32_t selFx, selFy, selFz;
          asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(selFx) : "r"(selNx));
          asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(selFy) : "r"(selNy));
          asm volatile("xor.b32 %0, %1, 0x22;" : "=r"(selFz) : "r"(selNz));
          const float bT = bestT * 1.0002f;
          float       tn[4];
          bool        h[4];
#pragma unroll
          for(int k = 0; k < 4; k++)
          {
            const float a = fmaxf(fmaxf(__fmaf_rn(gridPlane(w[3 * k + 0], selNx), id.x, nood.x), __fmaf_rn(gridPlane(w[3 * k + 1], selNy), id.y, nood.y)),
                                  __fmaf_rn(gridPlane(w[3 * k + 2], selNz), id.z, nood.z));
            const float b = fminf(fminf(__fmaf_rn(gridPlane(w[3 * k + 0], selFx), id.x, nood.x), __fmaf_rn(gridPlane(w[3 * k + 1], selFy), id.y, nood.y)),
                                  __fmaf_rn(gridPlane(w[3 * k + 2], selFz), id.z, nood.z));
            // max(a, 0) <= min(b, bestT) * 1.0002  <=>  a <= x && x >= 0 with x = min(b * 1.0002, bestT * 1.0002)   (bestT >= 0)
            const float x = fminf(b * 1.0002f, bT);
            h[k]  = (a <= x) && (x >= 0.0f);
            tn[k] = h[k] ? a : kInfinity;
          }
          // nearest hit child (empty slots repeat child 0 behind an inverted box, so they need no test of their own)
          const bool     s01 = tn[1] < tn[0], s23 = tn[3] < tn[2];
          const float    t01 = s01 ? tn[1] : tn[0], t23 = s23 ? tn[3] : tn[2];
          const uint32_t r01 = s01 ? w[13] : w[12], r23 = s23 ? w[15] : w[14];
          const bool     sab = t23 < t01;
          const uint32_t nearRef = sab ? r23 : r01;
          if(fminf(t01, t23) < kInfinity)
          {
            // enter the nearest child; the other hit children go on the stack.  Fast path: the slot above the top is always
            // written and the pointer moves only for a hit -- no branches.  Up to three pushes, and the store of the fourth
            // iteration lands one slot above them: slots sp .. sp + 3 must all be inside the shared array
            if(sp + 4 <= kStackShared)
            {
#pragma unroll
              for(int k = 0; k < 4; k++)
              {
                sstack[sp * kTraceBlock] = w[12 + k];
                sp += (h[k] && w[12 + k] != nearRef) ? 1 : 0;
              }
            }
            else
            {
#pragma unroll
              for(int k = 0; k < 4; k++)
                if(h[k] && w[12 + k] != nearRef)
                  push(w[12 + k]);
            }
            node = nearRef;
          }
          else
            node = pop();
          if((node & kLeafBit) && leaf == 0)
          {
            leaf = node;
            node = pop();
          }
          if(__popc(__ballot_sync(__activemask(), leaf == 0)) <= 8)
            break;
        }

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
            if(!(t > 0.0f && (t < bestT || (t == bestT && bestTri != ~0u && triId < bestTri))))
              continue;
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, __float_as_uint(q2.w), u, v, __float_as_uint(L.rayD[slot].w)))
              continue;
            bestT = t; bestU = u; bestV = v; bestTri = triId;
            if(MODE != MODE_CLOSEST)
            {
              node = kDone;
              sp   = 0;
              break;
            }
          }
          if(node & kLeafBit)
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

        if(!exhausted && __popc(__activemask()) < kRefetchThreshold)
          break;
      }

      if(node == kDone && leaf == 0)
      {
        if(COUNT && L.cost && MODE != MODE_ANY_PROBE)
        {
          // heat-map debug mode: the nanoseconds this lane held the ray (waiting for the other lanes of its warp included, as in the
          // reference's per-invocation clock), charged to its path (one closest and one shadow ray per path and launch)
          const uint32_t path = (MODE == MODE_SHADOW_RESOLVE) ? __float_as_uint(L.shC[slot].w) : slot;
          L.cost[path].w += (float)(deviceNs() - rayBase);
        }
        if(MODE == MODE_CLOSEST)
          stS(&L.hit[slot], (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(bestT, bestU, bestV, __uint_as_float(bestTri)));
        else if(MODE == MODE_ANY_PROBE)
          L.hit[slot] = (bestTri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(0.f, 0.f, 0.f, __uint_as_float(0u));
        else if(bestTri == ~0u)
        {
          const float4   c    = ldS(&L.shC[slot]);
          const uint32_t path = __float_as_uint(c.w);
          float4         rad  = ldS(&L.rad[path]);
          rad.x += c.x; rad.y += c.y; rad.z += c.z;
          stS(&L.rad[path], rad);
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
  __shared__ uint32_t s_stack[(kStackShared > 0 ? kStackShared : 1) * (kStackShared > 0 ? kTraceBlock : 1)];
  uint32_t* const sstack = &s_stack[threadIdx.x];
  uint32_t        lstack[kStackLocal];
  const uint32_t  n    = *L.count;
  if(MODE == MODE_SHADOW_RESOLVE && L.stats && blockIdx.x == 0 && threadIdx.x == 0)
  {
    atomicAdd(&L.stats->shadowRays, (unsigned long long)n);  // AnyHit() calls of this bounce (round 1 counted them in a kernel of their own)
    if(L.primary)
      atomicAdd(&L.stats->shadowRaysPrimary, (unsigned long long)n);
  }
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
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, __float_as_uint(q2.w), u, v, seed))
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
  __shared__ uint32_t   s_stack[(kStackShared > 0 ? kStackShared : 1) * kTraceBlock];
  __shared__ WarpShared s_ws[kTraceBlock / 32];
  uint32_t* const sstack = &s_stack[threadIdx.x];
  WarpShared&     ws     = s_ws[threadIdx.x >> 5];
  uint32_t        lstack[kStackLocal];
  const uint32_t  n    = *L.count;
  if(MODE == MODE_SHADOW_RESOLVE && L.stats && blockIdx.x == 0 && threadIdx.x == 0)
  {
    atomicAdd(&L.stats->shadowRays, (unsigned long long)n);  // AnyHit() calls of this bounce (round 1 counted them in a kernel of their own)
    if(L.primary)
      atomicAdd(&L.stats->shadowRaysPrimary, (unsigned long long)n);
  }
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
            if(!(flags & TRI_OPAQUE) && !alphaTestPasses(sc, triId, __float_as_uint(q2.w), u, v, rseed))
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

template <int MODE, bool COUNT, int VARIANT, int FMT>
static void launchTraceV(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  static int grid = 0;
  if(!grid)
  {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, traceKernel<MODE, COUNT, VARIANT, FMT>, kTraceBlock, 0);
    grid = numSMs * (perSM < 1 ? 1 : perSM);
  }
  traceKernel<MODE, COUNT, VARIANT, FMT><<<grid, kTraceBlock, 0, s>>>(scene, L);
}
// Node format read by the default traversal kernel (VKRT_NODE_FORMAT): "ce" centre / half-extent fp32 (default, 64 bytes),
// "lohi" min / max fp32 (64 bytes), "q32" 16-bit grid boxes (32 bytes).  The experimental loop variants only exist for "lohi".
int traceNodeFormat()
{
  static int fmt = -1;
  if(fmt < 0)
  {
    const char* e = getenv("VKRT_NODE_FORMAT");
    const char* v = getenv("VKRT_TRACE_VARIANT");
    fmt           = FMT_CE;
    if(e && !strcmp(e, "lohi")) fmt = FMT_LOHI;
    if(e && !strcmp(e, "q32")) fmt = FMT_Q32;
    if(v && atoi(v) != 8) fmt = FMT_LOHI;
  }
  return fmt;
}
int traceNodeBytes(int bvhWidth) { return (bvhWidth == 4) ? 64 : (traceNodeFormat() == FMT_Q32 ? 32 : 64); }
template <int MODE, bool COUNT>
static void launchTraceWide(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  static int grid = 0;
  if(!grid)
  {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, traceKernelWide<MODE, COUNT>, kTraceBlock, 0);
    grid = numSMs * (perSM < 1 ? 1 : perSM);
  }
  traceKernelWide<MODE, COUNT><<<grid, kTraceBlock, 0, s>>>(scene, L);
}
// Tree width for a scene (VKRT_BVH_WIDTH = 2 | 4 | auto, default auto): the binary tree while the BVH fits the L2, the 4-wide
// tree (half as many dependent fetches per ray) once it is HBM resident -- measured: equal on the 263 k-triangle scene,
// +9 % on the 10 M-triangle one.  The experimental loop variants only exist for the binary tree.
int traceBvhWidth(uint32_t numTris)
{
  static int mode = -1;  // 0 auto, 2, 4
  if(mode < 0)
  {
    const char* e = getenv("VKRT_BVH_WIDTH");
    const char* v = getenv("VKRT_TRACE_VARIANT");
    mode          = (e && atoi(e) == 4) ? 4 : ((e && atoi(e) == 2) ? 2 : 0);
    if(v && atoi(v) != 8) mode = 2;
  }
  if(mode == 0)
    return numTris > 1500000u ? 4 : 2;
  return mode;
}
template <int MODE, bool COUNT>
static void launchTrace(cudaStream_t s, const SceneView& scene, const TraceLaunch& L, int numSMs)
{
  if(scene.bvhWidth == 4)
  {
    launchTraceWide<MODE, COUNT>(s, scene, L, numSMs);
    return;
  }
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
    case 1: launchTraceV<MODE, COUNT, 1, FMT_LOHI>(s, scene, L, numSMs); break;
    case 4: launchTraceV<MODE, COUNT, 4, FMT_LOHI>(s, scene, L, numSMs); break;
    case 8:
      if(traceNodeFormat() == FMT_Q32)
        launchTraceV<MODE, COUNT, 8, FMT_Q32>(s, scene, L, numSMs);
      else if(traceNodeFormat() == FMT_CE)
        launchTraceV<MODE, COUNT, 8, FMT_CE>(s, scene, L, numSMs);
      else
        launchTraceV<MODE, COUNT, 8, FMT_LOHI>(s, scene, L, numSMs);
      break;
    case 16: launchTraceV<MODE, COUNT, 16, FMT_LOHI>(s, scene, L, numSMs); break;
    default: launchTraceV<MODE, COUNT, 0, FMT_LOHI>(s, scene, L, numSMs); break;
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
