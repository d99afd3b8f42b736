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
