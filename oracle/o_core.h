// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or imported by the product.
//
// o_core.h: integer-exact pieces of the reference restated for the CPU:
//   RNG            shaders/random.glsl:34-102
//   oct encoding   shaders/compress.glsl:60-180
//   misc math      shaders/common.glsl:27-113
#pragma once
#include "o_math.h"

namespace orc {

// ---- shaders/random.glsl ---------------------------------------------------
// random.glsl:34-48  16 rounds of TEA
inline uint32_t tea(uint32_t val0, uint32_t val1)
{
  uint32_t v0 = val0, v1 = val1, s0 = 0;
  for(uint32_t n = 0; n < 16; n++)
  {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
// random.glsl:50-53
inline uint32_t initRandom(uint32_t resx, uint32_t sx, uint32_t sy, uint32_t frame) { return tea(sy * resx + sx, frame); }
// random.glsl:59-65  PCG RXS-M-XS, state advanced in place
inline uint32_t pcg(uint32_t& state)
{
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state         = prev;
  return (word >> 22u) ^ word;
}
// random.glsl:81-92
inline void pcg3d(uint32_t v[3])
{
  v[0] = v[0] * 1664525u + 1013904223u;
  v[1] = v[1] * 1664525u + 1013904223u;
  v[2] = v[2] * 1664525u + 1013904223u;
  v[0] += v[1] * v[2];
  v[1] += v[2] * v[0];
  v[2] += v[0] * v[1];
  v[0] ^= v[0] >> 16u;
  v[1] ^= v[1] >> 16u;
  v[2] ^= v[2] >> 16u;
  v[0] += v[1] * v[2];
  v[1] += v[2] * v[0];
  v[2] += v[0] * v[1];
}
// random.glsl:98-102  float in [0,1)
inline float rnd(uint32_t& seed)
{
  uint32_t r = pcg(seed);
  return uintBitsToFloat(0x3f800000u | (r >> 9)) - 1.0f;
}

// Order-independent replacement for the per-candidate draw of HitTest
// (shaders/traceray_rq.glsl:98).  The reference draws rand(prd.seed) once per
// non-opaque candidate in driver traversal order, which is not reproducible
// across BVHs; this framework pins the draw to a hash of (ray seed, triangle id)
// and advances the path seed by exactly one pcg step per trace call when the
// scene holds non-opaque instances (DESIGN.md "alpha candidates").
inline float alphaDraw(uint32_t raySeed, uint32_t triId)
{
  uint32_t s = raySeed ^ (triId * 0x9e3779b9u + 0x85ebca6bu);
  uint32_t r = pcg(s);
  return uintBitsToFloat(0x3f800000u | (r >> 9)) - 1.0f;
}

// ---- shaders/compress.glsl -------------------------------------------------
// compress.glsl:74-98  (host-side roundEven of the reference, kept with its quirks)
inline float roundEvenHost(float x)
{
  int   Integer        = static_cast<int>(x);
  float IntegerPart    = static_cast<float>(Integer);
  float FractionalPart = (x - std::floor(x));
  if(FractionalPart > 0.5f || FractionalPart < 0.5f)
    return std::round(x);
  else if((Integer % 2) == 0)
    return IntegerPart;
  else if(x <= 0)
    return IntegerPart - 1;
  else
    return IntegerPart + 1;
}
// compress.glsl:60-72  glm::packUnorm4x8 as the reference's host fallback spells it
inline uint32_t packUnorm4x8(vec4 v)
{
  uint32_t b0 = (uint32_t)(unsigned char)std::round(std::fmin(std::fmax(v.x, 0.0f), 1.0f) * 255.f);
  uint32_t b1 = (uint32_t)(unsigned char)std::round(std::fmin(std::fmax(v.y, 0.0f), 1.0f) * 255.f);
  uint32_t b2 = (uint32_t)(unsigned char)std::round(std::fmin(std::fmax(v.z, 0.0f), 1.0f) * 255.f);
  uint32_t b3 = (uint32_t)(unsigned char)std::round(std::fmin(std::fmax(v.w, 0.0f), 1.0f) * 255.f);
  return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}
// GLSL unpackUnorm4x8: byte / 255.0
inline vec4 unpackUnorm4x8(uint32_t p)
{
  return {(float)(p & 0xffu) / 255.0f, (float)((p >> 8) & 0xffu) / 255.0f, (float)((p >> 16) & 0xffu) / 255.0f, (float)((p >> 24) & 0xffu) / 255.0f};
}
// compress.glsl:111-139
inline uint32_t compress_unit_vec(vec3 nv)
{
  const float C_Stack_Max = 3.402823466e+38f;
  if((nv.x < C_Stack_Max) && !std::isinf(nv.x))
  {
    const float d = 32767.0f / (std::fabs(nv.x) + std::fabs(nv.y) + std::fabs(nv.z));
    int         x = int(roundEvenHost(nv.x * d));
    int         y = int(roundEvenHost(nv.y * d));
    if(nv.z < 0.0f)
    {
      const int maskx = x >> 31;
      const int masky = y >> 31;
      const int tmp   = 32767 + maskx + masky;
      const int tmpx  = x;
      x               = (tmp - (y ^ masky)) ^ maskx;
      y               = (tmp - (tmpx ^ maskx)) ^ masky;
    }
    uint32_t packed = (uint32_t(y + 32767) << 16) | uint32_t(x + 32767);
    if(packed == ~0u)
      return ~0x1u;
    return packed;
  }
  return ~0u;
}
// compress.glsl:143-147
inline float short_to_floatm11(const int v)
{
  return (v >= 0) ? (uintBitsToFloat(0x3F800000u | (uint32_t(v) << 8)) - 1.0f) :
                    (uintBitsToFloat((0x80000000u | 0x3F800000u) | (uint32_t(-v) << 8)) + 1.0f);
}
// compress.glsl:149-180
inline vec3 decompress_unit_vec(uint32_t packed)
{
  if(packed != ~0u)
  {
    int       x     = int(packed & 0xFFFFu) - 32767;
    int       y     = int(packed >> 16) - 32767;
    const int maskx = x >> 31;
    const int masky = y >> 31;
    const int tmp0  = 32767 + maskx + masky;
    const int ymask = y ^ masky;
    const int tmp1  = tmp0 - (x ^ maskx);
    const int z     = tmp1 - ymask;
    float     zf;
    if(z < 0)
    {
      x  = (tmp0 - ymask) ^ maskx;
      y  = tmp1 ^ masky;
      zf = uintBitsToFloat((0x80000000u | 0x3F800000u) | (uint32_t(-z) << 8)) + 1.0f;
    }
    else
    {
      zf = uintBitsToFloat(0x3F800000u | (uint32_t(z) << 8)) - 1.0f;
    }
    return normalize(V3(short_to_floatm11(x), short_to_floatm11(y), zf));
  }
  return V3(3.402823466e+38f);
}

// ---- shaders/common.glsl ---------------------------------------------------
// common.glsl:67-74
inline vec2 GetSphericalUv(vec3 v)
{
  float gamma = std::asin(-v.y);
  float theta = std::atan2(v.z, v.x);
  return V2(theta * K_1_OVER_PI * 0.5f + 0.5f, gamma * K_1_OVER_PI + 0.5f);
}
// common.glsl:80-92
inline void CreateCoordinateSystem(vec3 N, vec3& Nt, vec3& Nb)
{
  Nt = normalize((std::fabs(N.z) > 0.99999f) ? V3(-N.x * N.y, 1.0f - N.y * N.y, -N.y * N.z) : V3(-N.x * N.z, -N.y * N.z, 1.0f - N.z * N.z));
  Nb = cross(Nt, N);
}
// common.glsl:98-113
inline vec3 OffsetRay(vec3 p, vec3 n)
{
  const float intScale   = 256.0f;
  const float floatScale = 1.0f / 65536.0f;
  const float origin     = 1.0f / 32.0f;
  int         ofx = (int)(intScale * n.x), ofy = (int)(intScale * n.y), ofz = (int)(intScale * n.z);
  vec3        p_i = V3(intBitsToFloat(floatBitsToInt(p.x) + ((p.x < 0) ? -ofx : ofx)),
                       intBitsToFloat(floatBitsToInt(p.y) + ((p.y < 0) ? -ofy : ofy)),
                       intBitsToFloat(floatBitsToInt(p.z) + ((p.z < 0) ? -ofz : ofz)));
  return V3(std::fabs(p.x) < origin ? p.x + floatScale * n.x : p_i.x,  //
            std::fabs(p.y) < origin ? p.y + floatScale * n.y : p_i.y,  //
            std::fabs(p.z) < origin ? p.z + floatScale * n.z : p_i.z);
}
// common.glsl:38-61  heat-map colours
inline float fadef(float low, float high, float value)
{
  float mid   = (low + high) * 0.5f;
  float range = (high - low) * 0.5f;
  float x     = 1.0f - clampf(std::fabs(mid - value) / range, 0.0f, 1.0f);
  return smoothstepf(0.0f, 1.0f, x);
}
inline vec3 temperature(float intensity)
{
  const vec3 blue = V3(0, 0, 1), cyan = V3(0, 1, 1), green = V3(0, 1, 0), yellow = V3(1, 1, 0), red = V3(1, 0, 0);
  return (fadef(-0.25f, 0.25f, intensity) * blue + fadef(0.0f, 0.5f, intensity) * cyan + fadef(0.25f, 0.75f, intensity) * green
          + fadef(0.5f, 1.0f, intensity) * yellow + smoothstepf(0.75f, 1.0f, intensity) * red);
}

}  // namespace orc
