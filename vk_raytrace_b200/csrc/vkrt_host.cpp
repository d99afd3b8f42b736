// vkrt_host.cpp -- host-side (CPU) halves of the reference's scene path that sit on the
// product side of the C ABI: they prepare the buffers the GPU kernels consume.
//   alias table          src/hdr_sampling.cpp:107-248  (HdrSampling::buildAliasmap / createEnvironmentAccel)
//   vertex packing       src/scene.cpp:224-246 + shaders/compress.glsl:60-139
//   camera matrices      src/scene.cpp:629-640
// Pure functions of their inputs; compiled with -ffp-contract=off so that the float
// results are reproducible (the alias table depends on float summation order).
#include <cmath>
#include <cstring>
#include <vector>
#include <numeric>
#include <algorithm>
#include "../../include/vkrt_b200.h"

namespace vkrt {

// ---- 4x4 inverse (double cofactors, rounded once to float) -------------------------------------
bool invertAffine(const float* m16, float* inv16)
{
  double a[16], inv[16];
  for(int i = 0; i < 16; i++) a[i] = m16[i];
  inv[0]  = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
  inv[4]  = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
  inv[8]  = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
  inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
  inv[1]  = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
  inv[5]  = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
  inv[9]  = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
  inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
  inv[2]  = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
  inv[6]  = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
  inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
  inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
  inv[3]  = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
  inv[7]  = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
  inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
  inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
  double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
  if(det == 0.0)
    return false;
  det = 1.0 / det;
  for(int i = 0; i < 16; i++) inv16[i] = (float)(inv[i] * det);
  return true;
}

// ---- environment importance sampling table -----------------------------------------------------
// Walker alias method over solid-angle-weighted max(R,G,B); float accumulation order kept.
static float buildAliasTable(const std::vector<float>& importance, EnvAccel* accel)
{
  const uint32_t n   = (uint32_t)importance.size();
  const float    sum = std::accumulate(importance.begin(), importance.end(), 0.f);
  const float    inverseAverage = (float)n / sum;
  for(uint32_t i = 0; i < n; ++i)
  {
    accel[i].q     = importance[i] * inverseAverage;
    accel[i].alias = i;
  }
  // below-average texels fill the table from the front, the others from the back
  std::vector<uint32_t> part(n);
  uint32_t              small = 0u, large = n;
  for(uint32_t i = 0; i < n; ++i)
  {
    if(accel[i].q < 1.f)
      part[small++] = i;
    else
      part[--large] = i;
  }
  for(small = 0; small < large && large < n; ++small)
  {
    const uint32_t lo = part[small];
    const uint32_t hi = part[large];
    accel[lo].alias   = hi;
    const float deficit = 1.f - accel[lo].q;
    accel[hi].q -= deficit;
    if(accel[hi].q < 1.0f)
      large++;
  }
  return sum;
}

void buildEnvAccelHost(const float* rgba, uint32_t w, uint32_t h, EnvAccel* out, float* integral, float* average)
{
  const size_t       n = (size_t)w * h;
  std::vector<float> importance(n);
  float              cosTheta0 = 1.0f;
  const float        stepPhi   = float(2.0 * M_PI) / float(w);
  const float        stepTheta = float(M_PI) / float(h);
  double             total     = 0;
  for(uint32_t y = 0; y < h; ++y)
  {
    const float theta1    = float(y + 1) * stepTheta;
    const float cosTheta1 = std::cos(theta1);
    const float area      = (cosTheta0 - cosTheta1) * stepPhi;
    cosTheta0             = cosTheta1;
    for(uint32_t x = 0; x < w; ++x)
    {
      const size_t i  = (size_t)y * w + x;
      const float* p  = rgba + i * 4;
      importance[i]   = area * std::max(p[0], std::max(p[1], p[2]));
      total += p[0] * 0.2126f + p[1] * 0.7152f + p[2] * 0.0722f;
    }
  }
  const float integ = buildAliasTable(importance, out);
  const float invIntegral = 1.0f / integ;
  for(size_t i = 0; i < n; ++i)
  {
    const float* p = rgba + i * 4;
    out[i].pdf     = std::max(p[0], std::max(p[1], p[2])) * invIntegral;
  }
  for(size_t i = 0; i < n; ++i)
    out[i].aliasPdf = out[out[i].alias].pdf;
  if(integral) *integral = integ;
  if(average) *average = static_cast<float>(total) / static_cast<float>(w * h);
}

// ---- vertex packing -----------------------------------------------------------------------------
static float roundEvenQuirk(float x)  // shaders/compress.glsl:74-98 (host fallback, odd negative halves included)
{
  const int   ip   = static_cast<int>(x);
  const float frac = x - std::floor(x);
  if(frac != 0.5f)
    return std::round(x);
  if((ip % 2) == 0)
    return (float)ip;
  return x <= 0 ? (float)ip - 1 : (float)ip + 1;
}
static uint32_t octEncode(float nx, float ny, float nz)  // shaders/compress.glsl:111-139
{
  if(!((nx < 3.402823466e+38f) && !std::isinf(nx)))
    return ~0u;
  const float d = 32767.0f / (std::fabs(nx) + std::fabs(ny) + std::fabs(nz));
  int         x = int(roundEvenQuirk(nx * d));
  int         y = int(roundEvenQuirk(ny * d));
  if(nz < 0.0f)
  {
    const int mx = x >> 31, my = y >> 31;
    const int t  = 32767 + mx + my;
    const int x0 = x;
    x            = (t - (y ^ my)) ^ mx;
    y            = (t - (x0 ^ mx)) ^ my;
  }
  const uint32_t packed = (uint32_t(y + 32767) << 16) | uint32_t(x + 32767);
  return packed == ~0u ? ~0x1u : packed;
}
static uint32_t unorm8(float v) { return (uint32_t)(unsigned char)std::round(std::min(std::max(v, 0.0f), 1.0f) * 255.f); }

}  // namespace vkrt

extern "C" {

int vkrt_pack_vertices(const float* positions, const float* normals, const float* tangents, const float* texcoords, const float* colors,
                       uint32_t n, VertexAttributes* out)
{
  if(!positions || !normals || !tangents || !texcoords || !colors || !out)
    return VKRT_ERR_INVALID;
  for(uint32_t i = 0; i < n; i++)
  {
    VertexAttributes v{};
    std::memcpy(v.position, positions + 3 * i, 12);
    v.normal      = vkrt::octEncode(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
    v.tangent     = vkrt::octEncode(tangents[4 * i], tangents[4 * i + 1], tangents[4 * i + 2]);
    v.texcoord[0] = texcoords[2 * i];
    uint32_t bits;
    std::memcpy(&bits, texcoords + 2 * i + 1, 4);
    if(tangents[4 * i + 3] > 0)
      bits |= 1u;  // H == +1
    else
      bits &= ~1u;
    std::memcpy(&v.texcoord[1], &bits, 4);
    const float* c = colors + 4 * i;
    v.color        = vkrt::unorm8(c[0]) | (vkrt::unorm8(c[1]) << 8) | (vkrt::unorm8(c[2]) << 16) | (vkrt::unorm8(c[3]) << 24);
    out[i]         = v;
  }
  return VKRT_OK;
}

int vkrt_make_camera(const float eye[3], const float center[3], const float up[3], float fov_degrees, float aspect, float aperture,
                     int32_t nb_lights, SceneCamera* out)
{
  if(!eye || !center || !up || !out)
    return VKRT_ERR_INVALID;
  auto sub = [](const float* a, const float* b, float* r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; };
  auto dot = [](const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  auto nrm = [&](float* a) { float l = std::sqrt(dot(a, a)); a[0] /= l; a[1] /= l; a[2] /= l; };
  auto crs = [](const float* a, const float* b, float* r) { r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; };
  float f[3], s[3], u[3];
  sub(center, eye, f);
  const float focal = std::sqrt(dot(f, f));
  nrm(f);
  crs(f, up, s);
  nrm(s);
  crs(s, f, u);
  float view[16] = {s[0], u[0], -f[0], 0.f, s[1], u[1], -f[1], 0.f, s[2], u[2], -f[2], 0.f, -dot(s, eye), -dot(u, eye), dot(f, eye), 1.f};
  const float zNear = 0.001f, zFar = 100000.0f;
  const float tanHalf = std::tan(fov_degrees * 0.01745329251994329576923690768489f / 2.0f);
  float proj[16] = {0};
  proj[0]  = 1.0f / (aspect * tanHalf);
  proj[5]  = -(1.0f / tanHalf);  // proj[1][1] *= -1
  proj[10] = zFar / (zNear - zFar);
  proj[11] = -1.0f;
  proj[14] = -(zFar * zNear) / (zFar - zNear);
  SceneCamera c{};
  if(!vkrt::invertAffine(view, c.viewInverse) || !vkrt::invertAffine(proj, c.projInverse))
    return VKRT_ERR_INVALID;
  c.focalDist = focal;
  c.aperture  = aperture;
  c.nbLights  = nb_lights;
  *out        = c;
  return VKRT_OK;
}

int vkrt_light_from_matrix(const float m[16], Light* l)
{
  if(!m || !l)
    return VKRT_ERR_INVALID;
  l->position[0] = m[12]; l->position[1] = m[13]; l->position[2] = m[14];
  l->direction[0] = -m[8]; l->direction[1] = -m[9]; l->direction[2] = -m[10];
  return VKRT_OK;
}

}  // extern "C"

// ---- host binned-SAH builder (VKRT_BUILD_SAH_HOST) ----------------------------------------------------
// The "prefer fast trace" analogue of the driver build the reference requests
// (VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT_KHR, src/accelstruct.cpp:125-126): a top-down
// binned surface-area-heuristic BVH2 over the world-space triangles the GPU flatten pass produced, emitted
// in the same 64-byte node format as the GPU LBVH.  Slower to build, fewer nodes per ray.
#include <atomic>
#include <future>
namespace vkrt {
namespace {
struct SahBuild
{
  const float*          tris;  // n x 12 floats (3 x float4)
  std::vector<float>    lo, hi, ce;  // n x 3
  std::vector<uint32_t> order;
  float*                nodes;  // 16 floats per node
  std::atomic<uint32_t> nextNode{1};
  std::atomic<int>      maxDepth{0};
};
struct Bounds
{
  float lo[3], hi[3];
  void  reset() { for(int k = 0; k < 3; k++) { lo[k] = 3e38f; hi[k] = -3e38f; } }
  void  grow(const float* l, const float* h) { for(int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); } }
  float area() const { float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return 2.f * (x * y + y * z + z * x); }
};
constexpr int   kBins = 32;
constexpr uint32_t kLeafMax = 4;

uint32_t sahRecurse(SahBuild& B, uint32_t b, uint32_t e, Bounds& outBox, int depth, int par)
{
  const uint32_t n = e - b;
  Bounds box, cbox;
  box.reset(); cbox.reset();
  for(uint32_t i = b; i < e; i++)
  {
    const uint32_t t = B.order[i];
    box.grow(&B.lo[3 * t], &B.hi[3 * t]);
    cbox.grow(&B.ce[3 * t], &B.ce[3 * t]);
  }
  outBox = box;
  auto makeLeaf = [&]() -> uint32_t {
    int d = B.maxDepth.load();
    while(depth > d && !B.maxDepth.compare_exchange_weak(d, depth)) {}
    return 0x80000000u | ((n - 1) << 28) | b;
  };
  if(n == 1)
    return makeLeaf();
  // binned SAH over the widest centroid axes
  int   bestAxis = -1, bestBin = -1;
  float bestCost = 3e38f;
  for(int a = 0; a < 3; a++)
  {
    const float cmin = cbox.lo[a], cmax = cbox.hi[a];
    if(!(cmax > cmin))
      continue;
    Bounds   bb[kBins];
    uint32_t bc[kBins] = {0};
    for(int k = 0; k < kBins; k++) bb[k].reset();
    const float scale = (float)kBins / (cmax - cmin);
    for(uint32_t i = b; i < e; i++)
    {
      const uint32_t t = B.order[i];
      int k = std::min(kBins - 1, (int)((B.ce[3 * t + a] - cmin) * scale));
      bc[k]++;
      bb[k].grow(&B.lo[3 * t], &B.hi[3 * t]);
    }
    float    rA[kBins];
    uint32_t rC[kBins];
    Bounds   acc;
    acc.reset();
    uint32_t c = 0;
    for(int k = kBins - 1; k > 0; k--)
    {
      if(bc[k]) acc.grow(bb[k].lo, bb[k].hi);
      c += bc[k];
      rA[k] = c ? acc.area() : 0.f;
      rC[k] = c;
    }
    acc.reset();
    c = 0;
    for(int k = 0; k < kBins - 1; k++)
    {
      if(bc[k]) acc.grow(bb[k].lo, bb[k].hi);
      c += bc[k];
      if(c == 0 || rC[k + 1] == 0)
        continue;
      const float cost = acc.area() * (float)c + rA[k + 1] * (float)rC[k + 1];
      if(cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = k; }
    }
  }
  const float leafCost = box.area() * (float)n;
  if(n <= kLeafMax && (bestAxis < 0 || leafCost <= 1.0f * box.area() + bestCost))
    return makeLeaf();
  uint32_t mid;
  if(bestAxis >= 0)
  {
    const float cmin = cbox.lo[bestAxis], scale = (float)kBins / (cbox.hi[bestAxis] - cmin);
    auto it = std::partition(B.order.begin() + b, B.order.begin() + e, [&](uint32_t t) {
      return std::min(kBins - 1, (int)((B.ce[3 * t + bestAxis] - cmin) * scale)) <= bestBin;
    });
    mid = (uint32_t)(it - B.order.begin());
    if(mid == b || mid == e)
      mid = b + n / 2;
  }
  else
  {
    if(n <= 8)
      return makeLeaf();  // coincident centroids: up to 8 triangles fit a leaf
    mid = b + n / 2;
  }
  const uint32_t idx = B.nextNode.fetch_add(1);
  Bounds         lb, rb;
  uint32_t       lc, rc;
  if(par > 0 && n > 20000)
  {
    auto fut = std::async(std::launch::async, [&]() { return sahRecurse(B, b, mid, lb, depth + 1, par - 1); });
    rc       = sahRecurse(B, mid, e, rb, depth + 1, par - 1);
    lc       = fut.get();
  }
  else
  {
    lc = sahRecurse(B, b, mid, lb, depth + 1, 0);
    rc = sahRecurse(B, mid, e, rb, depth + 1, 0);
  }
  float* N = B.nodes + 16ull * idx;
  N[0] = lb.lo[0]; N[1] = lb.hi[0]; N[2] = lb.lo[1]; N[3] = lb.hi[1];
  N[4] = rb.lo[0]; N[5] = rb.hi[0]; N[6] = rb.lo[1]; N[7] = rb.hi[1];
  N[8] = lb.lo[2]; N[9] = lb.hi[2]; N[10] = rb.lo[2]; N[11] = rb.hi[2];
  std::memcpy(&N[12], &lc, 4);
  std::memcpy(&N[13], &rc, 4);
  N[14] = 0.f; N[15] = 0.f;
  return idx;
}
}  // namespace

// tris: n x 12 floats in any order.  Outputs nodes (16 floats each) and the permutation to apply to `tris`.
bool buildSahHost(const float* tris, uint32_t n, std::vector<float>& nodesOut, std::vector<uint32_t>& orderOut, int& maxDepthOut)
{
  if(n == 0)
    return false;
  SahBuild B;
  B.tris = tris;
  B.lo.resize(3ull * n); B.hi.resize(3ull * n); B.ce.resize(3ull * n);
  B.order.resize(n);
  for(uint32_t i = 0; i < n; i++)
  {
    const float* t = tris + 12ull * i;
    for(int k = 0; k < 3; k++)
    {
      const float l = std::min(t[k], std::min(t[4 + k], t[8 + k])), h = std::max(t[k], std::max(t[4 + k], t[8 + k]));
      B.lo[3ull * i + k] = l; B.hi[3ull * i + k] = h; B.ce[3ull * i + k] = (l + h) * 0.5f;
    }
    B.order[i] = i;
  }
  nodesOut.assign(16ull * std::max<uint32_t>(n, 2), 0.f);
  B.nodes = nodesOut.data();
  // node 0 is the root and must be an inner node: build the two halves of the top split into it
  Bounds   rootBox;
  B.nextNode = 0;
  uint32_t root = sahRecurse(B, 0, n, rootBox, 0, 3);
  if(root & 0x80000000u)
  {
    // the whole scene fits one leaf: make a root whose two children are that leaf
    float* N = B.nodes;
    N[0] = rootBox.lo[0]; N[1] = rootBox.hi[0]; N[2] = rootBox.lo[1]; N[3] = rootBox.hi[1];
    N[4] = N[0]; N[5] = N[1]; N[6] = N[2]; N[7] = N[3];
    N[8] = rootBox.lo[2]; N[9] = rootBox.hi[2]; N[10] = N[8]; N[11] = N[9];
    std::memcpy(&N[12], &root, 4);
    std::memcpy(&N[13], &root, 4);
    B.nextNode = 1;
  }
  nodesOut.resize(16ull * B.nextNode.load());
  orderOut    = std::move(B.order);
  maxDepthOut = B.maxDepth.load();
  return true;
}
}  // namespace vkrt
