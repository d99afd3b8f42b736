// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or imported by the product.
//
// o_bsdf.h: CPU restatement of the two BSDF models of the reference.
//   Disney principled  shaders/pbr_disney.glsl:66-599
//   glTF PBR           shaders/pbr_gltf.glsl:29-554
// State / Material mirror shaders/globals.glsl:72-122.
#pragma once
#include "o_core.h"

namespace orc {

struct Material
{
  vec3  albedo;
  float specular;
  vec3  emission;
  float anisotropy;
  float metallic;
  float roughness;
  float subsurface;
  float specularTint;
  float sheen;
  vec3  sheenTint;
  float clearcoat;
  float clearcoatRoughness;
  float transmission;
  float ior;
  vec3  attenuationColor;
  float attenuationDistance;
  float ax, ay;
  vec3  f0;
  float alpha;
  bool  unlit;
  bool  thinwalled;
};

struct State
{
  float    eta;
  vec3     position, normal, ffnormal, tangent, bitangent;
  vec2     texCoord;
  bool     isSubsurface;
  uint32_t matID;
  Material mat;
};

const float PI_D     = 3.14159265358979323f;  // globals.glsl:26 "PI"
const float TWO_PI_D = 6.28318530717958648f;  // globals.glsl:27

// ---------------------------------------------------------------- Disney helpers
// pbr_disney.glsl:66-80
inline vec3 ImportanceSampleGTR1(float rgh, float r1, float r2)
{
  float a        = std::fmax(0.001f, rgh);
  float a2       = a * a;
  float phi      = r1 * TWO_PI_D;
  float cosTheta = std::sqrt((1.0f - std::pow(a2, 1.0f - r1)) / (1.0f - a2));
  float sinTheta = clampf(std::sqrt(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
  float sinPhi   = std::sin(phi);
  float cosPhi   = std::cos(phi);
  return V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
// pbr_disney.glsl:84-93
inline vec3 ImportanceSampleGTR2_aniso(float ax, float ay, float r1, float r2)
{
  float phi      = r1 * TWO_PI_D;
  float sinPhi   = ay * std::sin(phi);
  float cosPhi   = ax * std::cos(phi);
  float tanTheta = std::sqrt(r2 / (1 - r2));
  return V3(tanTheta * cosPhi, tanTheta * sinPhi, 1.0f);
}
// pbr_disney.glsl:97-109
inline vec3 ImportanceSampleGTR2(float rgh, float r1, float r2)
{
  float a        = std::fmax(0.001f, rgh);
  float phi      = r1 * TWO_PI_D;
  float cosTheta = std::sqrt((1.0f - r2) / (1.0f + (a * a - 1.0f) * r2));
  float sinTheta = clampf(std::sqrt(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
  float sinPhi   = std::sin(phi);
  float cosPhi   = std::cos(phi);
  return V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
// pbr_disney.glsl:113-118
inline float SchlickFresnel(float u)
{
  float m  = clampf(1.0f - u, 0.0f, 1.0f);
  float m2 = m * m;
  return m2 * m2 * m;
}
// pbr_disney.glsl:122-136
inline float DielectricFresnel(float cos_theta_i, float eta)
{
  float sinThetaTSq = eta * eta * (1.0f - cos_theta_i * cos_theta_i);
  if(sinThetaTSq > 1.0f)
    return 1.0f;
  float cos_theta_t = std::sqrt(std::fmax(1.0f - sinThetaTSq, 0.0f));
  float rs          = (eta * cos_theta_t - cos_theta_i) / (eta * cos_theta_t + cos_theta_i);
  float rp          = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
  return 0.5f * (rs * rs + rp * rp);
}
// pbr_disney.glsl:140-147
inline float GTR1(float NdotH, float a)
{
  if(a >= 1.0f)
    return K_1_OVER_PI;
  float a2 = a * a;
  float t  = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
  return (a2 - 1.0f) / (PI_D * std::log(a2) * t);
}
// pbr_disney.glsl:151-156
inline float GTR2(float NdotH, float a)
{
  float a2 = a * a;
  float t  = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
  return a2 / (PI_D * t * t);
}
// pbr_disney.glsl:160-166
inline float GTR2_aniso(float NdotH, float HdotX, float HdotY, float ax, float ay)
{
  float a = HdotX / ax;
  float b = HdotY / ay;
  float c = a * a + b * b + NdotH * NdotH;
  return 1.0f / (PI_D * ax * ay * c * c);
}
// pbr_disney.glsl:170-175
inline float SmithG_GGX(float NdotV, float alphaG)
{
  float a = alphaG * alphaG;
  float b = NdotV * NdotV;
  return 1.0f / (NdotV + std::sqrt(a + b - a * b));
}
// pbr_disney.glsl:179-185
inline float SmithG_GGX_aniso(float NdotV, float VdotX, float VdotY, float ax, float ay)
{
  float a = VdotX * ax;
  float b = VdotY * ay;
  float c = NdotV;
  return 1.0f / (NdotV + std::sqrt(a * a + b * b + c * c));
}
// pbr_disney.glsl:189-199
inline vec3 CosineSampleHemisphere(float r1, float r2)
{
  vec3  dir;
  float r   = std::sqrt(r1);
  float phi = TWO_PI_D * r2;
  dir.x     = r * std::cos(phi);
  dir.y     = r * std::sin(phi);
  dir.z     = std::sqrt(std::fmax(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
  return dir;
}
// pbr_disney.glsl:203-209
inline vec3 UniformSampleHemisphere(float r1, float r2)
{
  float r   = std::sqrt(std::fmax(0.0f, 1.0f - r1 * r1));
  float phi = TWO_PI_D * r2;
  return V3(r * std::cos(phi), r * std::sin(phi), r1);
}
// pbr_disney.glsl:224-229
inline float powerHeuristic(float a, float b)
{
  float t = a * a;
  return t / (b * b + t);
}

// pbr_disney.glsl:318-331
inline vec3 EvalDielectricReflection(const State& state, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return V3(0.0f);
  float F = DielectricFresnel(dot(V, H), state.eta);
  float D = GTR2(dot(N, H), state.mat.roughness);
  pdf     = D * dot(N, H) * F / (4.0f * dot(V, H));
  float G = SmithG_GGX(std::fabs(dot(N, L)), state.mat.roughness) * SmithG_GGX(dot(N, V), state.mat.roughness);
  return state.mat.albedo * F * D * G;
}
// pbr_disney.glsl:335-346
inline vec3 EvalDielectricRefraction(const State& state, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  float F         = DielectricFresnel(std::fabs(dot(V, H)), state.eta);
  float D         = GTR2(dot(N, H), state.mat.roughness);
  float denomSqrt = dot(L, H) * state.eta + dot(V, H);
  pdf             = D * dot(N, H) * (1.0f - F) * std::fabs(dot(L, H)) / (denomSqrt * denomSqrt);
  float G = SmithG_GGX(std::fabs(dot(N, L)), state.mat.roughness) * SmithG_GGX(dot(N, V), state.mat.roughness);
  return state.mat.albedo * (1.0f - F) * D * G * std::fabs(dot(V, H)) * std::fabs(dot(L, H)) * 4.0f * state.eta * state.eta
         / (denomSqrt * denomSqrt);
}
// pbr_disney.glsl:350-363
inline vec3 EvalSpecular(const State& state, vec3 Cspec0, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return V3(0.0f);
  float D  = GTR2_aniso(dot(N, H), dot(H, state.tangent), dot(H, state.bitangent), state.mat.ax, state.mat.ay);
  pdf      = D * dot(N, H) / (4.0f * dot(V, H));
  float FH = SchlickFresnel(dot(L, H));
  vec3  F  = mix(Cspec0, V3(1.0f), FH);
  float G  = SmithG_GGX_aniso(dot(N, L), dot(L, state.tangent), dot(L, state.bitangent), state.mat.ax, state.mat.ay);
  G *= SmithG_GGX_aniso(dot(N, V), dot(V, state.tangent), dot(V, state.bitangent), state.mat.ax, state.mat.ay);
  return F * D * G;
}
// pbr_disney.glsl:367-379
inline vec3 EvalClearcoat(const State& state, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return V3(0.0f);
  float D  = GTR1(dot(N, H), state.mat.clearcoatRoughness);
  pdf      = D * dot(N, H) / (4.0f * dot(V, H));
  float FH = SchlickFresnel(dot(L, H));
  float F  = mixf(0.04f, 1.0f, FH);
  float G  = SmithG_GGX(dot(N, L), 0.25f) * SmithG_GGX(dot(N, V), 0.25f);
  return V3(0.25f * state.mat.clearcoat * F * D * G);
}
// pbr_disney.glsl:383-397
inline vec3 EvalDiffuse(const State& state, vec3 Csheen, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return V3(0.0f);
  pdf          = dot(N, L) * (1.0f / PI_D);
  float FL     = SchlickFresnel(dot(N, L));
  float FV     = SchlickFresnel(dot(N, V));
  float FH     = SchlickFresnel(dot(L, H));
  float Fd90   = 0.5f + 2.0f * dot(L, H) * dot(L, H) * state.mat.roughness;
  float Fd     = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
  vec3  Fsheen = FH * state.mat.sheen * Csheen;
  return ((1.0f / PI_D) * Fd * (1.0f - state.mat.subsurface) * state.mat.albedo + Fsheen) * (1.0f - state.mat.metallic);
}
// pbr_disney.glsl:401-409
inline vec3 EvalSubsurface(const State& state, vec3 V, vec3 N, vec3 L, float& pdf)
{
  pdf      = (1.0f / TWO_PI_D);
  float FL = SchlickFresnel(std::fabs(dot(N, L)));
  float FV = SchlickFresnel(dot(N, V));
  float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
  return sqrt3(state.mat.albedo) * state.mat.subsurface * (1.0f / PI_D) * Fd * (1.0f - state.mat.metallic) * (1.0f - state.mat.transmission);
}

// pbr_disney.glsl:414-520   (state is inout: eta and isSubsurface are written)
inline vec3 DisneySample(State& state, vec3 V, vec3 N, vec3& L, float& pdf, uint32_t& seed)
{
  state.isSubsurface = false;
  pdf                = 0.0f;
  vec3 f             = V3(0.0f);

  float r1 = rnd(seed);
  float r2 = rnd(seed);

  float diffuseRatio = 0.5f * (1.0f - state.mat.metallic);
  float transWeight  = (1.0f - state.mat.metallic) * state.mat.transmission;

  vec3  Cdlin = state.mat.albedo;
  float Cdlum = 0.3f * Cdlin.x + 0.6f * Cdlin.y + 0.1f * Cdlin.z;

  vec3 Ctint  = Cdlum > 0.0f ? Cdlin / Cdlum : V3(1.0f);
  vec3 Cspec0 = mix(state.mat.specular * 0.08f * mix(V3(1.0f), Ctint, state.mat.specularTint), Cdlin, state.mat.metallic);
  vec3 Csheen = state.mat.sheenTint;

  if(rnd(seed) < transWeight)
  {
    vec3 H = ImportanceSampleGTR2(state.mat.roughness, r1, r2);
    H      = state.tangent * H.x + state.bitangent * H.y + N * H.z;

    vec3  R = reflect(-V, H);
    float F = DielectricFresnel(std::fabs(dot(R, H)), state.eta);

    if(state.mat.thinwalled)
    {
      if(dot(state.ffnormal, state.normal) < 0.0f)
        F = 0;
      state.eta = 1.001f;
    }

    if(rnd(seed) < F)
    {
      L = normalize(R);
      f = EvalDielectricReflection(state, V, N, L, H, pdf);
    }
    else
    {
      L = normalize(refract(-V, H, state.eta));
      f = EvalDielectricRefraction(state, V, N, L, H, pdf);
    }
    f *= transWeight;
    pdf *= transWeight;
  }
  else
  {
    if(rnd(seed) < diffuseRatio)
    {
      if(rnd(seed) < state.mat.subsurface)
      {
        L = UniformSampleHemisphere(r1, r2);
        L = state.tangent * L.x + state.bitangent * L.y - N * L.z;
        f = EvalSubsurface(state, V, N, L, pdf);
        pdf *= state.mat.subsurface * diffuseRatio;
        state.isSubsurface = true;
      }
      else
      {
        L      = CosineSampleHemisphere(r1, r2);
        L      = state.tangent * L.x + state.bitangent * L.y + N * L.z;
        vec3 H = normalize(L + V);
        f      = EvalDiffuse(state, Csheen, V, N, L, H, pdf);
        pdf *= (1.0f - state.mat.subsurface) * diffuseRatio;
      }
    }
    else
    {
      float primarySpecRatio = 1.0f / (1.0f + state.mat.clearcoat);
      if(rnd(seed) < primarySpecRatio)
      {
        vec3 H = ImportanceSampleGTR2_aniso(state.mat.ax, state.mat.ay, r1, r2);
        H      = state.tangent * H.x + state.bitangent * H.y + N * H.z;
        L      = normalize(reflect(-V, H));
        f      = EvalSpecular(state, Cspec0, V, N, L, H, pdf);
        pdf *= primarySpecRatio * (1.0f - diffuseRatio);
      }
      else
      {
        vec3 H = ImportanceSampleGTR1(state.mat.clearcoatRoughness, r1, r2);
        H      = state.tangent * H.x + state.bitangent * H.y + N * H.z;
        L      = normalize(reflect(-V, H));
        f      = EvalClearcoat(state, V, N, L, H, pdf);
        pdf *= (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
      }
    }
    f *= (1.0f - transWeight);
    pdf *= (1.0f - transWeight);
  }
  return f;
}

// pbr_disney.glsl:524-599
inline vec3 DisneyEval(const State& state, vec3 V, vec3 N, vec3 L, float& pdf)
{
  vec3 H;
  if(dot(N, L) < 0.0f)
    H = normalize(L * (1.0f / state.eta) + V);
  else
    H = normalize(L + V);
  if(dot(N, H) < 0.0f)
    H = -H;

  float diffuseRatio     = 0.5f * (1.0f - state.mat.metallic);
  float primarySpecRatio = 1.0f / (1.0f + state.mat.clearcoat);
  float transWeight      = (1.0f - state.mat.metallic) * state.mat.transmission;

  vec3  brdf = V3(0.0f), bsdf = V3(0.0f);
  float brdfPdf = 0.0f, bsdfPdf = 0.0f;

  if(transWeight > 0.0f)
  {
    if(dot(N, L) < 0.0f)
      bsdf = EvalDielectricRefraction(state, V, N, L, H, bsdfPdf);
    else
      bsdf = EvalDielectricReflection(state, V, N, L, H, bsdfPdf);
  }

  float m_pdf = 0.0f;
  if(transWeight < 1.0f)
  {
    if(dot(N, L) < 0.0f)
    {
      if(state.mat.subsurface > 0.0f)
      {
        brdf    = EvalSubsurface(state, V, N, L, m_pdf);
        brdfPdf = m_pdf * state.mat.subsurface * diffuseRatio;
      }
    }
    else
    {
      vec3  Cdlin  = state.mat.albedo;
      float Cdlum  = 0.3f * Cdlin.x + 0.6f * Cdlin.y + 0.1f * Cdlin.z;
      vec3  Ctint  = Cdlum > 0.0f ? Cdlin / Cdlum : V3(1.0f);
      vec3  Cspec0 = mix(state.mat.specular * 0.08f * mix(V3(1.0f), Ctint, state.mat.specularTint), Cdlin, state.mat.metallic);
      vec3  Csheen = state.mat.sheenTint;

      brdf += EvalDiffuse(state, Csheen, V, N, L, H, m_pdf);
      brdfPdf += m_pdf * (1.0f - state.mat.subsurface) * diffuseRatio;

      brdf += EvalSpecular(state, Cspec0, V, N, L, H, m_pdf);
      brdfPdf += m_pdf * primarySpecRatio * (1.0f - diffuseRatio);

      brdf += EvalClearcoat(state, V, N, L, H, m_pdf);
      brdfPdf += m_pdf * (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
    }
  }
  pdf = mixf(brdfPdf, bsdfPdf, transWeight);
  return mix(brdf, bsdf, transWeight);
}

// ------------------------------------------------------------------ glTF PBR
// pbr_gltf.glsl:36-44
inline float powf5(float x) { return std::pow(x, 5.0f); }
inline vec3  F_Schlick3(vec3 f0, vec3 f90, float VdotH) { return f0 + (f90 - f0) * powf5(clampf(1.0f - VdotH, 0.0f, 1.0f)); }
inline float F_Schlick1(float f0, float f90, float VdotH) { return f0 + (f90 - f0) * powf5(clampf(1.0f - VdotH, 0.0f, 1.0f)); }
// pbr_gltf.glsl:51-64
inline float V_GGX(float NdotL, float NdotV, float alphaRoughness)
{
  float a2   = alphaRoughness * alphaRoughness;
  float GGXV = NdotL * std::sqrt(NdotV * NdotV * (1.0f - a2) + a2);
  float GGXL = NdotV * std::sqrt(NdotL * NdotL * (1.0f - a2) + a2);
  float GGX  = GGXV + GGXL;
  if(GGX > 0.0f)
    return 0.5f / GGX;
  return 0.0f;
}
// pbr_gltf.glsl:68-74
inline float V_GGX_anisotropic(float NdotL, float NdotV, float BdotV, float TdotV, float TdotL, float BdotL, float at, float ab)
{
  float GGXV = NdotL * length(V3(at * TdotV, ab * BdotV, NdotV));
  float GGXL = NdotV * length(V3(at * TdotL, ab * BdotL, NdotL));
  float v    = 0.5f / (GGXV + GGXL);
  return clampf(v, 0.0f, 1.0f);
}
// pbr_gltf.glsl:96-101
inline float D_GGX(float NdotH, float alphaRoughness)
{
  float a2 = alphaRoughness * alphaRoughness;
  float f  = (NdotH * NdotH) * (a2 - 1.0f) + 1.0f;
  return a2 / (K_PI * f * f);
}
// pbr_gltf.glsl:106-112
inline float D_GGX_anisotropic(float NdotH, float TdotH, float BdotH, float at, float ab)
{
  float a2 = at * ab;
  vec3  f  = V3(ab * TdotH, at * BdotH, a2 * NdotH);
  float w2 = a2 / dot(f, f);
  return a2 * w2 * w2 / K_PI;
}
// pbr_gltf.glsl:130-136
inline vec3 BRDF_lambertian(vec3 diffuseColor, float metallic) { return (1.0f - metallic) * (diffuseColor / K_PI); }
// pbr_gltf.glsl:139-146
inline vec3 BRDF_specularGGX(vec3 f0, vec3 f90, float alphaRoughness, float VdotH, float NdotL, float NdotV, float NdotH)
{
  vec3  F = F_Schlick3(f0, f90, VdotH);
  float V = V_GGX(NdotL, NdotV, alphaRoughness);
  float D = D_GGX(NdotH, std::fmax(0.001f, alphaRoughness));
  return F * V * D;
}
// pbr_gltf.glsl:149-175
inline vec3 BRDF_specularAnisotropicGGX(vec3 f0, vec3 f90, float alphaRoughness, float VdotH, float NdotL, float NdotV, float NdotH,
                                        float BdotV, float TdotV, float TdotL, float BdotL, float TdotH, float BdotH, float anisotropy)
{
  float at = std::fmax(alphaRoughness * (1.0f + anisotropy), 0.00001f);
  float ab = std::fmax(alphaRoughness * (1.0f - anisotropy), 0.00001f);
  vec3  F  = F_Schlick3(f0, f90, VdotH);
  float V  = V_GGX_anisotropic(NdotL, NdotV, BdotV, TdotV, TdotL, BdotL, at, ab);
  float D  = D_GGX_anisotropic(NdotH, TdotH, BdotH, at, ab);
  return F * V * D;
}
// pbr_gltf.glsl:188-198
inline vec3 GgxSampling(float specularAlpha, float r1, float r2)
{
  float phi      = r1 * 2.0f * K_PI;
  float cosTheta = std::sqrt((1.0f - r2) / (1.0f + (specularAlpha * specularAlpha - 1.0f) * r2));
  float sinTheta = clampf(std::sqrt(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
  float sinPhi   = std::sin(phi);
  float cosPhi   = std::cos(phi);
  return V3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
// pbr_gltf.glsl:203-219
inline vec3 EvalDiffuseGltf(const State& state, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  pdf         = 0;
  float NdotV = dot(N, V);
  float NdotL = dot(N, L);
  if(NdotL < 0.0f || NdotV < 0.0f)
    return V3(0.0f);
  NdotL = clampf(NdotL, 0.001f, 1.0f);
  pdf   = NdotL * K_1_OVER_PI;
  return BRDF_lambertian(state.mat.albedo, state.mat.metallic);
}
// pbr_gltf.glsl:224-258
inline vec3 EvalAnisotropicSpecularGltf(const State& state, vec3 f0, vec3 f90, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  pdf         = 0;
  float NdotL = dot(N, L);
  if(NdotL < 0.0f)
    return V3(0.0f);
  vec3  T = state.tangent, B = state.bitangent;
  float TdotV = clampf(dot(T, V), 0, 1);
  float BdotV = clampf(dot(B, V), 0, 1);
  float TdotL = dot(T, L), BdotL = dot(B, L), TdotH = dot(T, H), BdotH = dot(B, H);
  float NdotH = dot(N, H), NdotV = dot(N, V), VdotH = dot(V, H), LdotH = dot(L, H);
  NdotL = clampf(NdotL, 0.001f, 1.0f);
  NdotV = clampf(std::fabs(NdotV), 0.001f, 1.0f);
  float at = std::fmax(state.mat.roughness * (1.0f + state.mat.anisotropy), 0.001f);
  float ab = std::fmax(state.mat.roughness * (1.0f - state.mat.anisotropy), 0.001f);
  pdf      = D_GGX_anisotropic(NdotH, TdotH, BdotH, at, ab) / (4.0f * LdotH);
  return BRDF_specularAnisotropicGGX(f0, f90, state.mat.roughness, VdotH, NdotL, NdotV, NdotH, BdotV, TdotV, TdotL, BdotL, TdotH, BdotH, state.mat.anisotropy);
}
// pbr_gltf.glsl:263-285
inline vec3 EvalSpecularGltf(const State& state, vec3 f0, vec3 f90, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  if(state.mat.anisotropy > 0)
    return EvalAnisotropicSpecularGltf(state, f0, f90, V, N, L, H, pdf);
  pdf         = 0;
  float NdotL = dot(N, L);
  if(NdotL < 0.0f)
    return V3(0.0f);
  float NdotV = dot(N, V);
  float NdotH = clampf(dot(N, H), 0, 1);
  float LdotH = clampf(dot(L, H), 0, 1);
  float VdotH = clampf(dot(V, H), 0, 1);
  NdotL = clampf(NdotL, 0.001f, 1.0f);
  NdotV = clampf(std::fabs(NdotV), 0.001f, 1.0f);
  pdf   = D_GGX(NdotH, state.mat.roughness) * NdotH / (4.0f * LdotH);
  return BRDF_specularGGX(f0, f90, state.mat.roughness, VdotH, NdotL, NdotV, NdotH);
}
// pbr_gltf.glsl:290-315
inline vec3 EvalClearcoatGltf(const State& state, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf)
{
  pdf         = 0;
  float NdotL = dot(N, L);
  if(NdotL < 0.0f)
    return V3(0.0f);
  float NdotH = dot(N, H), NdotV = dot(N, V), VdotH = dot(V, H), LdotH = dot(L, H);
  NdotL = clampf(NdotL, 0.001f, 1.0f);
  NdotV = clampf(std::fabs(NdotV), 0.001f, 1.0f);
  float clearcoat        = state.mat.clearcoat;
  float clearcoatFresnel = F_Schlick1(0.04f, 1, VdotH);
  float clearcoatAlpha   = state.mat.clearcoatRoughness * state.mat.clearcoatRoughness;
  float G                = V_GGX(NdotL, NdotV, clearcoatAlpha);
  float D                = D_GGX(NdotH, std::fmax(0.001f, clearcoatAlpha));
  pdf                    = D * NdotH / (4.0f * LdotH);
  return V3(clearcoatFresnel * D * G * clearcoat);
}
// pbr_gltf.glsl:341-345  (the dielectric refraction is a pass-through in the reference)
inline vec3 EvalDielectricRefractionGltf(const State& state, vec3 N, vec3 L, float& pdf)
{
  pdf = std::fabs(dot(N, L));
  return state.mat.albedo;
}
// pbr_gltf.glsl:365-434
inline vec3 PbrEval(const State& state, vec3 V, vec3 N, vec3 L, float& pdfOut)
{
  vec3 H;
  if(dot(N, L) < 0.0f)
    H = normalize(L * (1.0f / state.eta) + V);
  else
    H = normalize(L + V);
  if(dot(N, H) < 0.0f)
    H = -H;

  float transWeight = (1.0f - state.mat.metallic) * state.mat.transmission;
  vec3  brdf = V3(0.0f), bsdf = V3(0.0f);
  float brdfPdf = 0.0f, bsdfPdf = 0.0f;

  if(transWeight > 0.0f)
    bsdf = EvalDielectricRefractionGltf(state, N, L, bsdfPdf);

  if(transWeight < 1.0f && dot(N, L) > 0)
  {
    float pdf;
    float diffuseRatio     = 0.5f * (1.0f - state.mat.metallic);
    float specularRatio    = 1.0f - diffuseRatio;
    float primarySpecRatio = 1.0f / (1.0f + state.mat.clearcoat);

    vec3  specularCol = state.mat.f0;
    float reflectance = std::fmax(std::fmax(specularCol.x, specularCol.y), specularCol.z);
    vec3  f0          = specularCol;
    vec3  f90         = V3(clampf(reflectance * 50.0f, 0.0f, 1.0f));

    brdf += EvalDiffuseGltf(state, V, N, L, H, pdf);
    brdfPdf += pdf * diffuseRatio;

    brdf += EvalClearcoatGltf(state, V, N, L, H, pdf);
    brdfPdf += pdf * (1.0f - primarySpecRatio) * specularRatio;

    brdf += EvalSpecularGltf(state, f0, f90, V, N, L, H, pdf);
    brdfPdf += pdf * primarySpecRatio * specularRatio;
  }
  pdfOut = mixf(brdfPdf, bsdfPdf, transWeight);
  return mix(brdf, bsdf, transWeight);
}
// pbr_gltf.glsl:439-554   (state is `in`: no write-back)
inline vec3 PbrSample(const State& state, vec3 V, vec3 N, vec3& L, float& pdf, uint32_t& seed)
{
  pdf       = 0.0f;
  vec3 brdf = V3(0.0f);

  float probability   = rnd(seed);
  float diffuseRatio  = 0.5f * (1.0f - state.mat.metallic);
  float specularRatio = 1.0f - diffuseRatio;
  float transWeight   = (1.0f - state.mat.metallic) * state.mat.transmission;

  float r1 = rnd(seed);
  float r2 = rnd(seed);

  if(rnd(seed) < transWeight)
  {
    float eta = state.eta;
    float n1 = 1.0f, n2 = state.mat.ior;
    float R0          = (n1 - n2) / (n1 + n2);
    vec3  H           = GgxSampling(state.mat.roughness, r1, r2);
    H                 = state.tangent * H.x + state.bitangent * H.y + N * H.z;
    float VdotH       = dot(V, H);
    float F           = F_Schlick1(R0 * R0, 1.0f, VdotH);
    float discriminat = 1.0f - eta * eta * (1.0f - VdotH * VdotH);

    if(state.mat.thinwalled)
    {
      if(dot(state.ffnormal, state.normal) < 0.0f)
      {
        F           = 0;
        discriminat = 0;
      }
      eta = 1.00f;
    }

    if(discriminat < 0.0f || rnd(seed) < F)
    {
      L = normalize(reflect(-V, H));
    }
    else
    {
      L = normalize(refract(-V, H, eta));
      if(std::isnan(L.x) || std::isnan(L.y) || std::isnan(L.z))
        L = -V;
    }
    brdf = EvalDielectricRefractionGltf(state, N, L, pdf);
  }
  else
  {
    vec3  specularCol = state.mat.f0;
    float reflectance = std::fmax(std::fmax(specularCol.x, specularCol.y), specularCol.z);
    vec3  f0          = specularCol;
    vec3  f90         = V3(clampf(reflectance * 50.0f, 0.0f, 1.0f));
    vec3  T = state.tangent, B = state.bitangent;

    if(probability < diffuseRatio)
    {
      L      = CosineSampleHemisphere(r1, r2);
      L      = L.x * T + L.y * B + L.z * N;
      vec3 H = normalize(L + V);
      brdf   = EvalDiffuseGltf(state, V, N, L, H, pdf);
      pdf *= (1.0f - state.mat.subsurface) * diffuseRatio;
    }
    else
    {
      float primarySpecRatio = 1.0f / (1.0f + state.mat.clearcoat);
      float roughness;
      if(rnd(seed) < primarySpecRatio)
        roughness = state.mat.roughness;
      else
        roughness = state.mat.clearcoatRoughness;

      vec3 H = GgxSampling(roughness, r1, r2);
      H      = T * H.x + B * H.y + N * H.z;
      L      = reflect(-V, H);

      if(rnd(seed) < primarySpecRatio)
      {
        brdf = EvalSpecularGltf(state, f0, f90, V, N, L, H, pdf);
        pdf *= primarySpecRatio * specularRatio;
      }
      else
      {
        brdf = EvalClearcoatGltf(state, V, N, L, H, pdf);
        pdf *= (1.0f - primarySpecRatio) * specularRatio;
      }
    }
    brdf *= (1.0f - transWeight);
    pdf *= (1.0f - transWeight);
  }
  return brdf;
}

}  // namespace orc
