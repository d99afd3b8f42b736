// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or imported by the product.
//
// o_scene.h: scene container, texture sampling and the ray/scene queries that the
// reference delegates to the Vulkan driver.
//   instance flags        src/accelstruct.cpp:132-157
//   ClosestHit / AnyHit   shaders/traceray_rq.glsl:108-185
//   HitTest (alpha)       shaders/traceray_rq.glsl:32-102
// The driver's BVH is unspecified; what is pinned here (DESIGN.md "traversal
// semantics") is the *result*: nearest Moller-Trumbore hit over all world-space
// triangles, interval (0, tmax) exclusive, ties broken by the lower flattened
// triangle id, back-face culling per the instance flags.  A brute-force path and a
// binned-SAH BVH path must agree bit for bit (tests/test_oracle.py).
#pragma once
#include <vector>
#include <cstdio>
#include "o_core.h"
#include "../include/vkrt_host_device.h"

namespace orc {

enum TriFlags : uint32_t { TRI_OPAQUE = 1u, TRI_CULL_DISABLE = 2u, TRI_FLIP_FACING = 4u };

struct PrimMesh
{
  std::vector<VertexAttributes> vtx;
  std::vector<uint32_t>         idx;
  int32_t                       material;
};
struct Node
{
  float    world[16];
  float    worldInv[16];
  uint32_t primMesh;
  uint32_t triOffset;
  uint32_t flags;
};
struct Texture
{
  std::vector<uint8_t> rgba;
  uint32_t             w, h;
  int                  linear;  // mag filter (LOD 0 => magnification)
  int                  wrapS, wrapT;  // 0 repeat, 1 clamp, 2 mirror
};
struct Tri
{
  vec3     v0, v1, v2;
  uint32_t node, prim, flags;
};
struct Hit
{
  float    t, u, v;
  uint32_t tri;  // flattened triangle id, ~0u = miss
};
struct BvhNode
{
  float    lo[3], hi[3];
  uint32_t left, right;  // inner: children; leaf: left = first, right = count | 0x80000000
};

struct Scene
{
  std::vector<PrimMesh>          meshes;
  std::vector<Node>              nodes;
  std::vector<GltfShadeMaterial> materials;
  std::vector<Light>             lights;
  std::vector<Texture>           textures;
  std::vector<float>             envPixels;  // RGBA32F
  uint32_t                       envW = 0, envH = 0;
  std::vector<EnvAccel>          envAccel;
  SceneCamera                    camera{};
  SunAndSky                      sunsky{};
  // committed
  std::vector<Tri>      tris;
  std::vector<uint32_t> triOrder;  // BVH leaf order -> tri id
  std::vector<BvhNode>  bvh;
  bool                  hasNonOpaque = false;
  bool                  useBvh       = true;
};

// ---- texture sampling ---------------------------------------------------------
// Vulkan LOD-0 sampling: unnormalised coord = u*size - 0.5, full-precision weights.
inline int wrapCoord(int i, int n, int mode)
{
  if(mode == 1)
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
  if(mode == 2)
  {
    int period = 2 * n;
    int m      = i % period;
    if(m < 0)
      m += period;
    return m < n ? m : period - 1 - m;
  }
  int m = i % n;
  return m < 0 ? m + n : m;
}
inline vec4 texel8(const Texture& t, int x, int y)
{
  const uint8_t* p = &t.rgba[(size_t(y) * t.w + x) * 4];
  return {p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f};
}
inline float sanitizeUv(float x)
{
  if(!(x == x) || std::fabs(x) > 1e9f)
    return 0.0f;
  return x;
}
inline vec4 sampleTexture(const Texture& t, vec2 uv)
{
  float u = sanitizeUv(uv.x), v = sanitizeUv(uv.y);
  if(!t.linear)
  {
    int x = wrapCoord((int)std::floor(u * (float)t.w), t.w, t.wrapS);
    int y = wrapCoord((int)std::floor(v * (float)t.h), t.h, t.wrapT);
    return texel8(t, x, y);
  }
  float fx = u * (float)t.w - 0.5f, fy = v * (float)t.h - 0.5f;
  float x0f = std::floor(fx), y0f = std::floor(fy);
  float ax = fx - x0f, ay = fy - y0f;
  int   x0 = wrapCoord((int)x0f, t.w, t.wrapS), x1 = wrapCoord((int)x0f + 1, t.w, t.wrapS);
  int   y0 = wrapCoord((int)y0f, t.h, t.wrapT), y1 = wrapCoord((int)y0f + 1, t.h, t.wrapT);
  vec4  c00 = texel8(t, x0, y0), c10 = texel8(t, x1, y0), c01 = texel8(t, x0, y1), c11 = texel8(t, x1, y1);
  vec4  top = c00 * (1.0f - ax) + c10 * ax;
  vec4  bot = c01 * (1.0f - ax) + c11 * ax;
  return top * (1.0f - ay) + bot * ay;
}
// Environment map: LINEAR, U repeat, V clamp-to-edge (src/hdr_sampling.cpp:68-75).
inline vec3 sampleEnv(const Scene& s, vec2 uv)
{
  float u = sanitizeUv(uv.x), v = sanitizeUv(uv.y);
  int   w = (int)s.envW, h = (int)s.envH;
  float fx = u * (float)w - 0.5f, fy = v * (float)h - 0.5f;
  float x0f = std::floor(fx), y0f = std::floor(fy);
  float ax = fx - x0f, ay = fy - y0f;
  int   x0 = wrapCoord((int)x0f, w, 0), x1 = wrapCoord((int)x0f + 1, w, 0);
  int   y0 = wrapCoord((int)y0f, h, 1), y1 = wrapCoord((int)y0f + 1, h, 1);
  auto  px = [&](int x, int y) { const float* p = &s.envPixels[(size_t(y) * w + x) * 4]; return V3(p[0], p[1], p[2]); };
  vec3  top = px(x0, y0) * (1.0f - ax) + px(x1, y0) * ax;
  vec3  bot = px(x0, y1) * (1.0f - ax) + px(x1, y1) * ax;
  return top * (1.0f - ay) + bot * ay;
}

// ---- commit: flatten instances into world space -----------------------------------
inline float det3(const float* m)
{
  return m[0] * (m[5] * m[10] - m[9] * m[6]) - m[4] * (m[1] * m[10] - m[9] * m[2]) + m[8] * (m[1] * m[6] - m[5] * m[2]);
}
void buildBvh(Scene& s);

inline void commitScene(Scene& s)
{
  s.tris.clear();
  s.hasNonOpaque = false;
  for(size_t n = 0; n < s.nodes.size(); n++)
  {
    Node&                    nd  = s.nodes[n];
    const PrimMesh&          pm  = s.meshes[nd.primMesh];
    const GltfShadeMaterial& mat = s.materials[std::max(0, pm.material)];
    uint32_t                 fl  = 0;
    // src/accelstruct.cpp:145-149
    if(mat.alphaMode == 0 || (mat.pbrBaseColorFactor[3] == 1.0f && mat.pbrBaseColorTexture == -1))
      fl |= TRI_OPAQUE;
    if(mat.doubleSided == 1)
      fl |= TRI_CULL_DISABLE;
    // Vulkan decides facing in object space; a mirroring instance transform flips
    // the world-space winding, so remember to flip it back.
    if(det3(nd.world) < 0.0f)
      fl |= TRI_FLIP_FACING;
    if(!(fl & TRI_OPAQUE))
      s.hasNonOpaque = true;
    nd.flags     = fl;
    nd.triOffset = (uint32_t)s.tris.size();
    invert4(nd.world, nd.worldInv);
    for(size_t p = 0; p + 2 < pm.idx.size(); p += 3)
    {
      Tri          t;
      const float* a = pm.vtx[pm.idx[p + 0]].position;
      const float* b = pm.vtx[pm.idx[p + 1]].position;
      const float* c = pm.vtx[pm.idx[p + 2]].position;
      t.v0           = xformPoint(nd.world, V3(a[0], a[1], a[2]));
      t.v1           = xformPoint(nd.world, V3(b[0], b[1], b[2]));
      t.v2           = xformPoint(nd.world, V3(c[0], c[1], c[2]));
      t.node         = (uint32_t)n;
      t.prim         = (uint32_t)(p / 3);
      t.flags        = fl;
      s.tris.push_back(t);
    }
  }
  buildBvh(s);
}

// ---- alpha test: shaders/traceray_rq.glsl:32-102 ---------------------------------
inline bool hitTestAlpha(const Scene& s, const Tri& T, uint32_t triId, float bu, float bv, uint32_t raySeed)
{
  const Node&              nd  = s.nodes[T.node];
  const PrimMesh&          pm  = s.meshes[nd.primMesh];
  const GltfShadeMaterial& mat = s.materials[std::max(0, pm.material)];
  float                    baseColorAlpha = mat.pbrBaseColorFactor[3];
  if(mat.pbrBaseColorTexture > -1)
  {
    const VertexAttributes& a0 = pm.vtx[pm.idx[T.prim * 3 + 0]];
    const VertexAttributes& a1 = pm.vtx[pm.idx[T.prim * 3 + 1]];
    const VertexAttributes& a2 = pm.vtx[pm.idx[T.prim * 3 + 2]];
    const float             b0 = 1.0f - bu - bv;
    vec2 tc = V2(a0.texcoord[0], a0.texcoord[1]) * b0 + V2(a1.texcoord[0], a1.texcoord[1]) * bu + V2(a2.texcoord[0], a2.texcoord[1]) * bv;
    // (vec4(tc,1,1) * uvTransform).xy : row vector times matrix
    const float* M = mat.uvTransform;
    vec2         tt;
    tt.x = ((tc.x * M[0] + tc.y * M[1]) + M[2]) + M[3];
    tt.y = ((tc.x * M[4] + tc.y * M[5]) + M[6]) + M[7];
    baseColorAlpha *= sampleTexture(s.textures[mat.pbrBaseColorTexture], tt).w;
  }
  float opacity;
  if(mat.alphaMode == VKRT_ALPHA_MASK)
    opacity = baseColorAlpha > mat.alphaCutoff ? 1.0f : 0.0f;
  else
    opacity = baseColorAlpha;
  if(alphaDraw(raySeed, triId) > opacity)
    return false;
  return true;
}

// ---- triangle test (pinned operation order, no FMA) ------------------------------
inline bool intersectTri(vec3 o, vec3 d, const Tri& T, bool cullBack, float& t, float& u, float& v)
{
  vec3  e1  = T.v1 - T.v0;
  vec3  e2  = T.v2 - T.v0;
  vec3  p   = cross(d, e2);
  float det = dot(e1, p);
  if(cullBack && !(T.flags & TRI_CULL_DISABLE))
  {
    bool front = (T.flags & TRI_FLIP_FACING) ? (det < 0.0f) : (det > 0.0f);
    if(!front)
      return false;
  }
  if(det == 0.0f || !(det == det))
    return false;
  float inv = 1.0f / det;
  vec3  tv  = o - T.v0;
  u         = dot(tv, p) * inv;
  if(!(u >= 0.0f && u <= 1.0f))
    return false;
  vec3 q = cross(tv, e1);
  v      = dot(d, q) * inv;
  if(!(v >= 0.0f && (u + v) <= 1.0f))
    return false;
  t = dot(e2, q) * inv;
  return true;
}

struct TraceCounters
{
  uint64_t nodes = 0, tris = 0;
};

// Candidate acceptance shared by brute force and BVH traversal.
inline void considerTri(const Scene& s, uint32_t id, vec3 o, vec3 d, float tmax, uint32_t raySeed, Hit& best)
{
  const Tri& T = s.tris[id];
  float      t, u, v;
  if(!intersectTri(o, d, T, true, t, u, v))
    return;
  if(!(t > 0.0f && t < tmax))
    return;
  if(!(t < best.t || (t == best.t && id < best.tri)))
    return;
  if(!(T.flags & TRI_OPAQUE) && !hitTestAlpha(s, T, id, u, v, raySeed))
    return;
  best.t = t; best.u = u; best.v = v; best.tri = id;
}

// Conservative slab test; returns the entry distance through tnOut.
inline bool slab(const BvhNode& n, vec3 o, vec3 inv, float tbest, float& tnOut)
{
  float t0x = (n.lo[0] - o.x) * inv.x, t1x = (n.hi[0] - o.x) * inv.x;
  float t0y = (n.lo[1] - o.y) * inv.y, t1y = (n.hi[1] - o.y) * inv.y;
  float t0z = (n.lo[2] - o.z) * inv.z, t1z = (n.hi[2] - o.z) * inv.z;
  float tn  = std::fmax(std::fmax(std::fmin(t0x, t1x), std::fmin(t0y, t1y)), std::fmax(std::fmin(t0z, t1z), 0.0f));
  float tf  = std::fmin(std::fmin(std::fmax(t0x, t1x), std::fmax(t0y, t1y)), std::fmin(std::fmax(t0z, t1z), tbest));
  tnOut     = tn;
  return tn <= tf * 1.0002f;  // never reject a box the triangle test could hit
}

// closest != 0: nearest hit.  closest == 0: any hit (terminate on first accepted).
inline Hit traceRay(const Scene& s, vec3 o, vec3 d, float tmax, uint32_t raySeed, bool closest, TraceCounters* cnt = nullptr)
{
  Hit best{tmax, 0.f, 0.f, ~0u};
  if(!s.useBvh || s.bvh.empty())
  {
    for(uint32_t i = 0; i < s.tris.size(); i++)
    {
      considerTri(s, i, o, d, tmax, raySeed, best);
      if(!closest && best.tri != ~0u)
        break;
    }
    if(cnt) cnt->tris += s.tris.size();
    return best;
  }
  // zero direction components are nudged so that 0 * inf never produces a NaN slab
  auto     safeInv = [](float x) { return 1.0f / (std::fabs(x) < 1e-20f ? std::copysign(1e-20f, x) : x); };
  vec3     inv     = V3(safeInv(d.x), safeInv(d.y), safeInv(d.z));
  uint32_t stack[128];
  int      sp = 0;
  float    tn;
  if(cnt) cnt->nodes++;
  if(!slab(s.bvh[0], o, inv, best.t, tn))
    return best;
  stack[sp++] = 0;
  // The visiting order (near child first) only affects speed: acceptance is order independent.
  while(sp)
  {
    const BvhNode& n = s.bvh[stack[--sp]];
    if(n.right & 0x80000000u)
    {
      uint32_t cntTri = n.right & 0x7fffffffu;
      for(uint32_t i = 0; i < cntTri; i++)
      {
        considerTri(s, s.triOrder[n.left + i], o, d, tmax, raySeed, best);
        if(cnt) cnt->tris++;
      }
      if(!closest && best.tri != ~0u)
        return best;
      continue;
    }
    float tl, tr;
    bool  hl = slab(s.bvh[n.left], o, inv, best.t, tl);
    bool  hr = slab(s.bvh[n.right], o, inv, best.t, tr);
    if(cnt) cnt->nodes += 2;
    if(hl && hr)
    {
      if(tr < tl)
      {
        stack[sp++] = n.left;
        stack[sp++] = n.right;
      }
      else
      {
        stack[sp++] = n.right;
        stack[sp++] = n.left;
      }
    }
    else if(hl)
      stack[sp++] = n.left;
    else if(hr)
      stack[sp++] = n.right;
  }
  return best;
}

// ---- binned SAH BVH2 ---------------------------------------------------------
inline void buildBvh(Scene& s)
{
  const size_t N = s.tris.size();
  s.bvh.clear();
  s.triOrder.resize(N);
  if(N == 0)
    return;
  std::vector<vec3> lo(N), hi(N), ce(N);
  for(size_t i = 0; i < N; i++)
  {
    const Tri& T = s.tris[i];
    lo[i]        = V3(std::fmin(T.v0.x, std::fmin(T.v1.x, T.v2.x)), std::fmin(T.v0.y, std::fmin(T.v1.y, T.v2.y)), std::fmin(T.v0.z, std::fmin(T.v1.z, T.v2.z)));
    hi[i]        = V3(std::fmax(T.v0.x, std::fmax(T.v1.x, T.v2.x)), std::fmax(T.v0.y, std::fmax(T.v1.y, T.v2.y)), std::fmax(T.v0.z, std::fmax(T.v1.z, T.v2.z)));
    ce[i]        = (lo[i] + hi[i]) * 0.5f;
    s.triOrder[i] = (uint32_t)i;
  }
  struct Job { uint32_t node, first, count; };
  std::vector<Job> jobs;
  s.bvh.reserve(2 * N);
  s.bvh.push_back(BvhNode{});
  jobs.push_back({0, 0, (uint32_t)N});
  auto comp = [](vec3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); };
  while(!jobs.empty())
  {
    Job j = jobs.back();
    jobs.pop_back();
    vec3 blo = V3(1e30f), bhi = V3(-1e30f), clo = V3(1e30f), chi = V3(-1e30f);
    for(uint32_t i = j.first; i < j.first + j.count; i++)
    {
      uint32_t id = s.triOrder[i];
      blo = V3(std::fmin(blo.x, lo[id].x), std::fmin(blo.y, lo[id].y), std::fmin(blo.z, lo[id].z));
      bhi = V3(std::fmax(bhi.x, hi[id].x), std::fmax(bhi.y, hi[id].y), std::fmax(bhi.z, hi[id].z));
      clo = V3(std::fmin(clo.x, ce[id].x), std::fmin(clo.y, ce[id].y), std::fmin(clo.z, ce[id].z));
      chi = V3(std::fmax(chi.x, ce[id].x), std::fmax(chi.y, ce[id].y), std::fmax(chi.z, ce[id].z));
    }
    BvhNode nd;
    nd.lo[0] = blo.x; nd.lo[1] = blo.y; nd.lo[2] = blo.z;
    nd.hi[0] = bhi.x; nd.hi[1] = bhi.y; nd.hi[2] = bhi.z;
    bool makeLeaf = j.count <= 2;
    int  bestAxis = -1, bestSplit = -1;
    const int NB = 16;
    if(!makeLeaf)
    {
      float bestCost = 1e30f;
      for(int a = 0; a < 3; a++)
      {
        float cmin = comp(clo, a), cmax = comp(chi, a);
        if(!(cmax > cmin))
          continue;
        struct Bin { vec3 lo, hi; uint32_t n; } bins[NB];
        for(int b = 0; b < NB; b++) { bins[b].lo = V3(1e30f); bins[b].hi = V3(-1e30f); bins[b].n = 0; }
        float scale = (float)NB / (cmax - cmin);
        for(uint32_t i = j.first; i < j.first + j.count; i++)
        {
          uint32_t id = s.triOrder[i];
          int      b  = std::min(NB - 1, (int)((comp(ce[id], a) - cmin) * scale));
          bins[b].n++;
          bins[b].lo = V3(std::fmin(bins[b].lo.x, lo[id].x), std::fmin(bins[b].lo.y, lo[id].y), std::fmin(bins[b].lo.z, lo[id].z));
          bins[b].hi = V3(std::fmax(bins[b].hi.x, hi[id].x), std::fmax(bins[b].hi.y, hi[id].y), std::fmax(bins[b].hi.z, hi[id].z));
        }
        float    rArea[NB];
        uint32_t rCnt[NB];
        vec3     rl = V3(1e30f), rh = V3(-1e30f);
        uint32_t rc = 0;
        auto area = [](vec3 l, vec3 h) { vec3 e = h - l; return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x); };
        for(int b = NB - 1; b > 0; b--)
        {
          if(bins[b].n) { rl = V3(std::fmin(rl.x, bins[b].lo.x), std::fmin(rl.y, bins[b].lo.y), std::fmin(rl.z, bins[b].lo.z)); rh = V3(std::fmax(rh.x, bins[b].hi.x), std::fmax(rh.y, bins[b].hi.y), std::fmax(rh.z, bins[b].hi.z)); }
          rc += bins[b].n;
          rArea[b] = rc ? area(rl, rh) : 0.f;
          rCnt[b]  = rc;
        }
        vec3     ll = V3(1e30f), lh = V3(-1e30f);
        uint32_t lc = 0;
        for(int b = 0; b < NB - 1; b++)
        {
          if(bins[b].n) { ll = V3(std::fmin(ll.x, bins[b].lo.x), std::fmin(ll.y, bins[b].lo.y), std::fmin(ll.z, bins[b].lo.z)); lh = V3(std::fmax(lh.x, bins[b].hi.x), std::fmax(lh.y, bins[b].hi.y), std::fmax(lh.z, bins[b].hi.z)); }
          lc += bins[b].n;
          if(lc == 0 || rCnt[b + 1] == 0)
            continue;
          float cost = area(ll, lh) * lc + rArea[b + 1] * rCnt[b + 1];
          if(cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = b; }
        }
      }
      if(bestAxis < 0)
        makeLeaf = j.count <= 8;  // all centroids coincide
    }
    uint32_t mid = j.first + j.count / 2;
    if(!makeLeaf && bestAxis >= 0)
    {
      float cmin = comp(clo, bestAxis), cmax = comp(chi, bestAxis);
      float scale = (float)NB / (cmax - cmin);
      auto  it = std::partition(s.triOrder.begin() + j.first, s.triOrder.begin() + j.first + j.count, [&](uint32_t id) {
        int b = std::min(NB - 1, (int)((comp(ce[id], bestAxis) - cmin) * scale));
        return b <= bestSplit;
      });
      mid = (uint32_t)(it - s.triOrder.begin());
      if(mid == j.first || mid == j.first + j.count)
        mid = j.first + j.count / 2;
    }
    if(makeLeaf)
    {
      nd.left  = j.first;
      nd.right = j.count | 0x80000000u;
      s.bvh[j.node] = nd;
    }
    else
    {
      nd.left  = (uint32_t)s.bvh.size();
      nd.right = nd.left + 1;
      s.bvh[j.node] = nd;
      s.bvh.push_back(BvhNode{});
      s.bvh.push_back(BvhNode{});
      jobs.push_back({nd.left, j.first, mid - j.first});
      jobs.push_back({nd.right, mid, j.first + j.count - mid});
    }
  }
}

}  // namespace orc
