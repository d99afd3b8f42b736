// vkrt_math.cuh -- device-side vector math with GLSL semantics for the sm_100a kernels.
//
// Operation orders that GLSL leaves open are fixed (dot = (x*x'+y*y')+z*z',
// normalize = v * (1/sqrt(dot)), mix = a*(1-t)+b*t) and the translation units are
// compiled with -fmad=false, so arithmetic is IEEE-exact and independent of
// scheduling; fused multiply-adds are written out explicitly (__fmaf_rn) only
// where rounding does not feed shading results (BVH slab tests).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define VKRT_HD __host__ __device__ __forceinline__
#define VKRT_D __device__ __forceinline__

namespace vkrt {

constexpr float kPi        = 3.14159265358979323846f;
constexpr float kTwoPi     = 6.28318530717958648f;
constexpr float kInvPi     = 0.318309886183790671538f;
constexpr float kInfinity  = 1e32f;  // shaders/globals.glsl:29

VKRT_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
VKRT_HD float3 f3(float s) { return make_float3(s, s, s); }
VKRT_HD float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
VKRT_HD float2 f2(float x, float y) { return make_float2(x, y); }
VKRT_HD float4 f4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
VKRT_HD float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }

VKRT_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
VKRT_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
VKRT_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
VKRT_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
VKRT_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
VKRT_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
VKRT_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
VKRT_HD float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
VKRT_HD float3 operator-(float3 a, float s) { return f3(a.x - s, a.y - s, a.z - s); }
VKRT_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
VKRT_HD float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
VKRT_HD float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
VKRT_HD float3& operator*=(float3& a, float s) { a = a * s; return a; }
VKRT_HD float3& operator/=(float3& a, float s) { a = a / s; return a; }
VKRT_HD float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
VKRT_HD float2 operator*(float2 a, float s) { return f2(a.x * s, a.y * s); }
VKRT_HD float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
VKRT_HD float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
VKRT_HD float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

VKRT_HD float  dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
VKRT_HD float  dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
VKRT_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
VKRT_HD float  length(float3 a) { return sqrtf(dot(a, a)); }
#if defined(__CUDA_ARCH__) && defined(VKRT_COMPACT_MATH)
static __device__ __noinline__ float3 normalize(float3 a)
#else
VKRT_HD float3 normalize(float3 a)
#endif
{
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
VKRT_HD float  clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
VKRT_HD float  mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
VKRT_HD float3 mix(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }
VKRT_HD float3 mix(float3 a, float3 b, float3 t) { return a * (f3(1.0f) - t) + b * t; }
VKRT_HD float  smoothstepf(float e0, float e1, float x)
{
  float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
// Transcendentals behind one call site each: the CUDA libm bodies are 50-200 instructions (sinf/cosf carry a
// Payne-Hanek slow path), and the shade kernel uses dozens of them -- inlining every one blew the kernel up to
// 16k instructions and made it instruction-fetch bound (profiles/README.md).  Results are unchanged.
#if defined(__CUDA_ARCH__) && defined(VKRT_COMPACT_MATH)
#define VKRT_MATHFN static __device__ __noinline__
#else
#define VKRT_MATHFN VKRT_HD
#endif
#if defined(__CUDA_ARCH__) && defined(VKRT_FAST_MATH)
// shading only (vkrt_shade.cu): SFU versions, a few ulp from libm -- the same order as the libm-vs-GLSL differences the
// image tolerances already allow for; traversal and intersection never use these
VKRT_HD float m_sinf(float x) { return __sinf(x); }
VKRT_HD float m_cosf(float x) { return __cosf(x); }
VKRT_HD float m_powf(float x, float y) { return __powf(x, y); }
VKRT_HD float m_expf(float x) { return __expf(x); }
VKRT_HD float m_logf(float x) { return __logf(x); }
#else
VKRT_MATHFN float m_sinf(float x) { return sinf(x); }
VKRT_MATHFN float m_cosf(float x) { return cosf(x); }
VKRT_MATHFN float m_powf(float x, float y) { return powf(x, y); }
VKRT_MATHFN float m_expf(float x) { return expf(x); }
VKRT_MATHFN float m_logf(float x) { return logf(x); }
#endif
VKRT_MATHFN float m_acosf(float x) { return acosf(x); }
VKRT_MATHFN float m_asinf(float x) { return asinf(x); }
VKRT_MATHFN float m_atan2f(float y, float x) { return atan2f(y, x); }

VKRT_HD float3 clamp3(float3 v, float lo, float hi) { return f3(clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)); }
VKRT_HD float3 pow3(float3 v, float e) { return f3(m_powf(v.x, e), m_powf(v.y, e), m_powf(v.z, e)); }
VKRT_HD float3 exp3(float3 v) { return f3(m_expf(v.x), m_expf(v.y), m_expf(v.z)); }
VKRT_HD float3 log3(float3 v) { return f3(m_logf(v.x), m_logf(v.y), m_logf(v.z)); }
VKRT_HD float3 sqrt3(float3 v) { return f3(sqrtf(v.x), sqrtf(v.y), sqrtf(v.z)); }
VKRT_HD float3 floor3(float3 v) { return f3(floorf(v.x), floorf(v.y), floorf(v.z)); }
VKRT_HD float  maxcomp(float3 v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }

VKRT_HD float3 reflect(float3 I, float3 N) { return I - N * (2.0f * dot(N, I)); }
VKRT_HD float3 refract(float3 I, float3 N, float eta)
{
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if(k < 0.0f)
    return f3(0.0f);
  return I * eta - N * (eta * d + sqrtf(k));
}

#ifdef __CUDA_ARCH__
VKRT_D uint32_t f2u(float f) { return __float_as_uint(f); }
VKRT_D int32_t  f2i(float f) { return __float_as_int(f); }
VKRT_D float    u2f(uint32_t u) { return __uint_as_float(u); }
VKRT_D float    i2f(int32_t i) { return __int_as_float(i); }
#else
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline int32_t  f2i(float f) { int32_t u; memcpy(&u, &f, 4); return u; }
inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float    i2f(int32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif

// 3x4 affine transform stored row-major: r[row*4 + col], 12 floats.
// point:  M * (p,1)   order ((m0*x + m1*y) + (m2*z + m3)) -- same as the oracle's xformPoint.
VKRT_HD float3 xformPoint(const float* r, float3 p)
{
  return f3((r[0] * p.x + r[1] * p.y) + (r[2] * p.z + r[3]), (r[4] * p.x + r[5] * p.y) + (r[6] * p.z + r[7]),
            (r[8] * p.x + r[9] * p.y) + (r[10] * p.z + r[11]));
}
VKRT_HD float3 xformVector(const float* r, float3 v)
{
  return f3((r[0] * v.x + r[1] * v.y) + r[2] * v.z, (r[4] * v.x + r[5] * v.y) + r[6] * v.z, (r[8] * v.x + r[9] * v.y) + r[10] * v.z);
}
// v * M (row vector times the 3x3 part): transpose multiply, used with worldToObject for normals.
VKRT_HD float3 xformNormalT(const float* r, float3 v)
{
  return f3((v.x * r[0] + v.y * r[4]) + v.z * r[8], (v.x * r[1] + v.y * r[5]) + v.z * r[9], (v.x * r[2] + v.y * r[6]) + v.z * r[10]);
}
// Column-major 4x4 (glm) times vec4, order ((c0*x + c1*y) + (c2*z + c3*w)).
VKRT_HD float4 mulMV(const float* m, float4 v)
{
  return f4((m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * v.w), (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * v.w),
            (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w), (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w));
}

// ---------------------------------------------------------------------------------
// RNG: shaders/random.glsl:34-102 (integer exact)
VKRT_HD uint32_t tea(uint32_t val0, uint32_t val1)
{
  uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
  for(int n = 0; n < 16; n++)
  {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
VKRT_HD uint32_t pcg(uint32_t& state)
{
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state         = prev;
  return (word >> 22u) ^ word;
}
VKRT_HD float rnd(uint32_t& seed)
{
  uint32_t r = pcg(seed);
  return u2f(0x3f800000u | (r >> 9)) - 1.0f;
}
// Order-independent alpha draw (DESIGN.md "alpha candidates"; replaces the sequential
// rand(prd.seed) of shaders/traceray_rq.glsl:98).
VKRT_HD float alphaDraw(uint32_t raySeed, uint32_t triId)
{
  uint32_t s = raySeed ^ (triId * 0x9e3779b9u + 0x85ebca6bu);
  uint32_t r = pcg(s);
  return u2f(0x3f800000u | (r >> 9)) - 1.0f;
}
VKRT_HD void pcg3d(uint32_t& x, uint32_t& y, uint32_t& z)
{
  x = x * 1664525u + 1013904223u;
  y = y * 1664525u + 1013904223u;
  z = z * 1664525u + 1013904223u;
  x += y * z; y += z * x; z += x * y;
  x ^= x >> 16u; y ^= y >> 16u; z ^= z >> 16u;
  x += y * z; y += z * x; z += x * y;
}

// ---------------------------------------------------------------------------------
// oct decode: shaders/compress.glsl:143-180
VKRT_HD float short_to_floatm11(int v)
{
  return (v >= 0) ? (u2f(0x3F800000u | (uint32_t(v) << 8)) - 1.0f) : (u2f(0xBF800000u | (uint32_t(-v) << 8)) + 1.0f);
}
#if defined(__CUDA_ARCH__) && defined(VKRT_COMPACT_MATH)
static __device__ __noinline__ float3 decompress_unit_vec(uint32_t packed)
#else
VKRT_HD float3 decompress_unit_vec(uint32_t packed)
#endif
{
  if(packed == ~0u)
    return f3(3.402823466e+38f);
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
    zf = u2f(0xBF800000u | (uint32_t(-z) << 8)) + 1.0f;
  }
  else
  {
    zf = u2f(0x3F800000u | (uint32_t(z) << 8)) - 1.0f;
  }
  return normalize(f3(short_to_floatm11(x), short_to_floatm11(y), zf));
}
VKRT_HD float4 unpackUnorm4x8(uint32_t p)
{
  return f4((float)(p & 0xffu) / 255.0f, (float)((p >> 8) & 0xffu) / 255.0f, (float)((p >> 16) & 0xffu) / 255.0f, (float)(p >> 24) / 255.0f);
}

// ---------------------------------------------------------------------------------
// shaders/common.glsl:67-113
VKRT_HD float2 GetSphericalUv(float3 v)
{
  float gamma = m_asinf(-v.y);
  float theta = m_atan2f(v.z, v.x);
  return f2(theta * kInvPi * 0.5f + 0.5f, gamma * kInvPi + 0.5f);
}
VKRT_HD void CreateCoordinateSystem(float3 N, float3& Nt, float3& Nb)
{
  Nt = normalize((fabsf(N.z) > 0.99999f) ? f3(-N.x * N.y, 1.0f - N.y * N.y, -N.y * N.z) : f3(-N.x * N.z, -N.y * N.z, 1.0f - N.z * N.z));
  Nb = cross(Nt, N);
}
VKRT_HD float3 OffsetRay(float3 p, float3 n)
{
  const float intScale   = 256.0f;
  const float floatScale = 1.0f / 65536.0f;
  const float origin     = 1.0f / 32.0f;
  int         ox = (int)(intScale * n.x), oy = (int)(intScale * n.y), oz = (int)(intScale * n.z);
  float3      pi = f3(i2f(f2i(p.x) + ((p.x < 0) ? -ox : ox)), i2f(f2i(p.y) + ((p.y < 0) ? -oy : oy)), i2f(f2i(p.z) + ((p.z < 0) ? -oz : oz)));
  return f3(fabsf(p.x) < origin ? p.x + floatScale * n.x : pi.x, fabsf(p.y) < origin ? p.y + floatScale * n.y : pi.y,
            fabsf(p.z) < origin ? p.z + floatScale * n.z : pi.z);
}

}  // namespace vkrt
