// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or imported by the product.
//
// o_math.h: GLSL vector semantics used by the CPU restatement of the reference
// shaders.  Evaluation orders that GLSL leaves to the implementation are pinned
// here (and mirrored by the CUDA kernels): dot = (x*x' + y*y') + z*z',
// normalize = v * (1/sqrt(dot(v,v))), mix = a*(1-t) + b*t.
// Compile with -ffp-contract=off so no FMA is formed behind our back.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

inline vec2 V2(float x, float y) { return {x, y}; }
inline vec3 V3(float x, float y, float z) { return {x, y, z}; }
inline vec3 V3(float s) { return {s, s, s}; }
inline vec4 V4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline vec4 V4(vec3 v, float w) { return {v.x, v.y, v.z, w}; }
inline vec3 xyz(vec4 v) { return {v.x, v.y, v.z}; }

inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }
inline vec2 operator*(float s, vec2 a) { return {a.x * s, a.y * s}; }
inline vec2 operator/(vec2 a, vec2 b) { return {a.x / b.x, a.y / b.y}; }

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }

inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator*(vec4 a, vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }

inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline vec3  cross(vec3 a, vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3  normalize(vec3 a) { float inv = 1.0f / std::sqrt(dot(a, a)); return a * inv; }

inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  mix(vec3 a, vec3 b, vec3 t) { return a * (V3(1.0f) - t) + b * t; }
inline float smoothstepf(float e0, float e1, float x)
{
  float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
inline vec3 clamp3(vec3 v, float lo, float hi) { return {clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)}; }
inline vec3 pow3(vec3 v, float e) { return {std::pow(v.x, e), std::pow(v.y, e), std::pow(v.z, e)}; }
inline vec3 pow3(vec3 v, vec3 e) { return {std::pow(v.x, e.x), std::pow(v.y, e.y), std::pow(v.z, e.z)}; }
inline vec3 exp3(vec3 v) { return {std::exp(v.x), std::exp(v.y), std::exp(v.z)}; }
inline vec3 log3(vec3 v) { return {std::log(v.x), std::log(v.y), std::log(v.z)}; }
inline vec3 sqrt3(vec3 v) { return {std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)}; }
inline vec3 max3(vec3 a, vec3 b) { return {std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)}; }
inline vec3 floor3(vec3 v) { return {std::floor(v.x), std::floor(v.y), std::floor(v.z)}; }

// GLSL reflect / refract
inline vec3 reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }
inline vec3 refract(vec3 I, vec3 N, float eta)
{
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if(k < 0.0f)
    return V3(0.0f);
  return I * eta - N * (eta * d + std::sqrt(k));
}

inline uint32_t floatBitsToUint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int32_t  floatBitsToInt(float f) { int32_t u; std::memcpy(&u, &f, 4); return u; }
inline float    uintBitsToFloat(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float    intBitsToFloat(int32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Column-major 4x4 (glm layout): m[c*4 + r].
struct mat4 { float m[16]; };
// M * vec4(p, w): pinned order ((c0*x + c1*y) + (c2*z + c3*w))
inline vec4 mulMV(const float* m, vec4 v)
{
  vec4 r;
  r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * v.w);
  r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * v.w);
  r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w);
  r.w = (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w);
  return r;
}
// 4x3 object->world applied to a point: rows of the 3x4 part of a column-major mat4.
inline vec3 xformPoint(const float* m, vec3 p)
{
  return {(m[0] * p.x + m[4] * p.y) + (m[8] * p.z + m[12]), (m[1] * p.x + m[5] * p.y) + (m[9] * p.z + m[13]),
          (m[2] * p.x + m[6] * p.y) + (m[10] * p.z + m[14])};
}
// M * vec4(v, 0)
inline vec3 xformVector(const float* m, vec3 v)
{
  return {(m[0] * v.x + m[4] * v.y) + m[8] * v.z, (m[1] * v.x + m[5] * v.y) + m[9] * v.z, (m[2] * v.x + m[6] * v.y) + m[10] * v.z};
}
// vec3(v * mat4x3) in GLSL (row-vector times matrix) == transpose(M3x3) * v : used with worldToObject.
inline vec3 xformNormalT(const float* m, vec3 v)
{
  return {(v.x * m[0] + v.y * m[1]) + v.z * m[2], (v.x * m[4] + v.y * m[5]) + v.z * m[6], (v.x * m[8] + v.y * m[9]) + v.z * m[10]};
}

// General 4x4 inverse (cofactors), used for worldToObject and the camera matrices.
inline bool invert4(const float* m, float* out)
{
  double inv[16], a[16];
  for(int i = 0; i < 16; i++) a[i] = m[i];
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
  for(int i = 0; i < 16; i++) out[i] = (float)(inv[i] * det);
  return true;
}

const float K_PI        = 3.14159265358979323846f;
const float K_TWO_PI    = 6.28318530717958648f;
const float K_1_OVER_PI = 0.318309886183790671538f;
const float INFINITY_RT  = 1e32f;  // shaders/globals.glsl:29

}  // namespace orc
