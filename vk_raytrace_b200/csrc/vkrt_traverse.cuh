// vkrt_traverse.cuh -- one-ray-per-thread BVH2 traversal for the fused tail kernel (vkrt_shade.cu: the thin late bounces of a path
// run trace -> shade -> shadow in one kernel instead of three launches per bounce).  Same result semantics as the persistent
// kernels of vkrt_trace.cu (DESIGN.md "traversal semantics"): nearest Moller-Trumbore hit with t in (0, tmax), ties to the lower
// triangle id, back faces culled unless the instance is double sided, non-opaque candidates pass the hashed alpha test; box
// culling is conservative, so the result is independent of visiting order -- the two implementations return identical bits.
//
// This header is compiled into both builds of vkrt_shade.cu, one of them with -prec-div=false: every division whose result can
// reach a hit record is written __fdiv_rn (IEEE, whatever the flags), and there is no square root on this path.
#pragma once
#include "vkrt_device.cuh"
#include "vkrt_texture.cuh"

namespace vkrt {

// the any-hit alpha test of vkrt_trace.cu (HitTest, shaders/traceray_rq.glsl:32-102), same arithmetic
static __device__ __noinline__ bool tailAlphaTestPasses(const SceneView& sc, uint32_t triId, uint32_t matId, float bu, float bv, uint32_t raySeed)
{
  const AlphaMat& am = sc.alphaMats[matId];
  const uint4*    ts = sc.triShade + 7ull * triId;
  const uint4     a1 = __ldg(ts + 1), b1 = __ldg(ts + 3), c1 = __ldg(ts + 5);
  float           baseColorAlpha = am.alphaFactor;
  if(am.hasTexture)
  {
    const float b0 = 1.0f - bu - bv;
    float2      tc = f2(__uint_as_float(a1.x), __uint_as_float(a1.y)) * b0 + f2(__uint_as_float(b1.x), __uint_as_float(b1.y)) * bu
                + f2(__uint_as_float(c1.x), __uint_as_float(c1.y)) * bv;
    tc             = applyUvTransform(am.uvRows, tc);
    baseColorAlpha *= sampleTexture(am.texture, tc).w;
  }
  float opacity = (am.alphaMode == VKRT_ALPHA_MASK) ? (baseColorAlpha > am.alphaCutoff ? 1.0f : 0.0f) : baseColorAlpha;
  return !(alphaDraw(raySeed, triId) > opacity);
}

struct TailHit
{
  float    t, u, v;
  uint32_t tri;  // ~0u: miss
};

template <bool ANY>
static __device__ __noinline__ TailHit traceSingle(const SceneView& sc, float3 o, float3 d, float tmax, uint32_t raySeed)
{
  TailHit best{tmax, 0.f, 0.f, ~0u};
  if(sc.numTris == 0)
    return best;
  const float4* __restrict__ nodes = sc.nodes;  // fp32 min/max boxes: always built, whatever tree the wavefront kernels walk
  const float4* __restrict__ tris  = sc.tris;
  auto inv = [](float x) { return __fdiv_rn(1.0f, fabsf(x) < 1e-20f ? copysignf(1e-20f, x) : x); };
  const float3 id = f3(inv(d.x), inv(d.y), inv(d.z));
  uint32_t     stack[128];
  int          sp   = 0;
  uint32_t     node = 0;
  const uint32_t kEnd = 0x7fffffffu;
  while(node != kEnd)
  {
    if(!(node & kLeafBit))
    {
      const float4 n0 = __ldg(&nodes[4ull * node]), n1 = __ldg(&nodes[4ull * node + 1]), n2 = __ldg(&nodes[4ull * node + 2]), n3 = __ldg(&nodes[4ull * node + 3]);
      const float a0x = (n0.x - o.x) * id.x, b0x = (n0.y - o.x) * id.x;
      const float a0y = (n0.z - o.y) * id.y, b0y = (n0.w - o.y) * id.y;
      const float a0z = (n2.x - o.z) * id.z, b0z = (n2.y - o.z) * id.z;
      const float a1x = (n1.x - o.x) * id.x, b1x = (n1.y - o.x) * id.x;
      const float a1y = (n1.z - o.y) * id.y, b1y = (n1.w - o.y) * id.y;
      const float a1z = (n2.z - o.z) * id.z, b1z = (n2.w - o.z) * id.z;
      const float tn0 = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fmaxf(fminf(a0z, b0z), 0.0f));
      const float tf0 = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fminf(fmaxf(a0z, b0z), best.t));
      const float tn1 = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fmaxf(fminf(a1z, b1z), 0.0f));
      const float tf1 = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fminf(fmaxf(a1z, b1z), best.t));
      const bool  h0 = tn0 <= tf0 * 1.0002f, h1 = tn1 <= tf1 * 1.0002f;  // conservative, as in traceKernel
      uint32_t    c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
      if(h0 && h1)
      {
        if(tn1 < tn0) { const uint32_t t = c0; c0 = c1; c1 = t; }
        if(sp < 128) stack[sp++] = c1;
        node = c0;
      }
      else if(h0)
        node = c0;
      else if(h1)
        node = c1;
      else
        node = sp ? stack[--sp] : kEnd;
      continue;
    }
    const uint32_t first = node & kLeafMask, cnt = ((node >> kLeafCntSh) & 7u) + 1u;
    node = sp ? stack[--sp] : kEnd;
    for(uint32_t k = 0; k < cnt; k++)
    {
      const float4 q0 = __ldg(&tris[3ull * (first + k) + 0]);
      const float4 q1 = __ldg(&tris[3ull * (first + k) + 1]);
      const float4 q2 = __ldg(&tris[3ull * (first + k) + 2]);
      const uint32_t flags = __float_as_uint(q0.w);
      const uint32_t triId = __float_as_uint(q1.w);
      // Moller-Trumbore, the pinned operation order of vkrt_trace.cu (no FMA: -fmad=false)
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
      const float  rdet = __fdiv_rn(1.0f, det);
      const float3 tv   = o - v0;
      const float  u    = dot(tv, p) * rdet;
      if(!(u >= 0.0f && u <= 1.0f))
        continue;
      const float3 q = cross(tv, e1);
      const float  v = dot(d, q) * rdet;
      if(!(v >= 0.0f && (u + v) <= 1.0f))
        continue;
      const float t = dot(e2, q) * rdet;
      if(!(t > 0.0f && (t < best.t || (t == best.t && best.tri != ~0u && triId < best.tri))))
        continue;
      if(!(flags & TRI_OPAQUE) && !tailAlphaTestPasses(sc, triId, __float_as_uint(q2.w), u, v, raySeed))
        continue;
      best.t = t; best.u = u; best.v = v; best.tri = triId;
      if(ANY)
        return best;  // TerminateOnFirstHit
    }
  }
  return best;
}

}  // namespace vkrt
