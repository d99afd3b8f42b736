// vkrt_bsdf.cuh -- device BSDF models of the shade kernel.
//   Disney principled BSDF   reference shaders/pbr_disney.glsl:66-599
//   glTF metallic-roughness  reference shaders/pbr_gltf.glsl:29-554
// The shading record (ShadeMat / ShadeFrame) is the register-resident equivalent of
// `State` + `Material` (shaders/globals.glsl:72-122); only fields that are read are kept.
#pragma once
#include "vkrt_math.cuh"

namespace vkrt {

struct ShadeMat
{
  float3 albedo, emission, sheenTint, attenuationColor, f0;
  float  metallic, roughness, transmission, ior, clearcoat, clearcoatRoughness;
  float  anisotropy, sheen, attenuationDistance, ax, ay, alpha;
  float  specular, specularTint, subsurface;
  bool   unlit, thinwalled;
};

struct ShadeFrame
{
  float3 normal, ffnormal, tangent, bitangent;
  float  eta;
  bool   isSubsurface;
};

// ----------------------------------------------------------------------------- shared lobes
VKRT_D float SchlickFresnel(float u)
{
  float m  = clampf(1.0f - u, 0.0f, 1.0f);
  float m2 = m * m;
  return m2 * m2 * m;
}
VKRT_D float DielectricFresnel(float cosThetaI, float eta)
{
  float sinThetaTSq = eta * eta * (1.0f - cosThetaI * cosThetaI);
  if(sinThetaTSq > 1.0f)
    return 1.0f;  // total internal reflection
  float cosThetaT = sqrtf(fmaxf(1.0f - sinThetaTSq, 0.0f));
  float rs        = (eta * cosThetaT - cosThetaI) / (eta * cosThetaT + cosThetaI);
  float rp        = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
  return 0.5f * (rs * rs + rp * rp);
}
VKRT_D float GTR1(float NdotH, float a)
{
  if(a >= 1.0f)
    return kInvPi;
  float a2 = a * a;
  float t  = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
  return (a2 - 1.0f) / (kPi * m_logf(a2) * t);
}
VKRT_D float GTR2(float NdotH, float a)
{
  float a2 = a * a;
  float t  = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
  return a2 / (kPi * t * t);
}
VKRT_D float GTR2_aniso(float NdotH, float HdotX, float HdotY, float ax, float ay)
{
  float a = HdotX / ax;
  float b = HdotY / ay;
  float c = a * a + b * b + NdotH * NdotH;
  return 1.0f / (kPi * ax * ay * c * c);
}
VKRT_D float SmithG_GGX(float NdotV, float alphaG)
{
  float a = alphaG * alphaG;
  float b = NdotV * NdotV;
  return 1.0f / (NdotV + sqrtf(a + b - a * b));
}
VKRT_D float SmithG_GGX_aniso(float NdotV, float VdotX, float VdotY, float ax, float ay)
{
  float a = VdotX * ax;
  float b = VdotY * ay;
  float c = NdotV;
  return 1.0f / (NdotV + sqrtf(a * a + b * b + c * c));
}
VKRT_D float powerHeuristic(float a, float b)
{
  float t = a * a;
  return t / (b * b + t);
}

// ----------------------------------------------------------------------------- sampling maps
VKRT_D float3 SampleGTR1(float rgh, float r1, float r2)
{
  float a        = fmaxf(0.001f, rgh);
  float a2       = a * a;
  float phi      = r1 * kTwoPi;
  float cosTheta = sqrtf((1.0f - m_powf(a2, 1.0f - r1)) / (1.0f - a2));
  float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
  float sinPhi = m_sinf(phi), cosPhi = m_cosf(phi);
  (void)r2;
  return f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
VKRT_D float3 SampleGTR2Aniso(float ax, float ay, float r1, float r2)
{
  float phi      = r1 * kTwoPi;
  float sinPhi   = ay * m_sinf(phi);
  float cosPhi   = ax * m_cosf(phi);
  float tanTheta = sqrtf(r2 / (1 - r2));
  return f3(tanTheta * cosPhi, tanTheta * sinPhi, 1.0f);
}
VKRT_D float3 SampleGTR2(float rgh, float r1, float r2)
{
  float a        = fmaxf(0.001f, rgh);
  float phi      = r1 * kTwoPi;
  float cosTheta = sqrtf((1.0f - r2) / (1.0f + (a * a - 1.0f) * r2));
  float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
  float sinPhi = m_sinf(phi), cosPhi = m_cosf(phi);
  return f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
VKRT_D float3 CosineSampleHemisphere(float r1, float r2)
{
  float  r   = sqrtf(r1);
  float  phi = kTwoPi * r2;
  float3 dir;
  dir.x = r * m_cosf(phi);
  dir.y = r * m_sinf(phi);
  dir.z = sqrtf(fmaxf(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
  return dir;
}
VKRT_D float3 UniformSampleHemisphere(float r1, float r2)
{
  float r   = sqrtf(fmaxf(0.0f, 1.0f - r1 * r1));
  float phi = kTwoPi * r2;
  return f3(r * m_cosf(phi), r * m_sinf(phi), r1);
}

// ----------------------------------------------------------------------------- Disney lobes
// pbr_disney.glsl:318-409
VKRT_D float3 DisneyDielectricReflection(const ShadeMat& m, float eta, float3 V, float3 N, float3 L, float3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return f3(0.0f);
  float F = DielectricFresnel(dot(V, H), eta);
  float D = GTR2(dot(N, H), m.roughness);
  pdf     = D * dot(N, H) * F / (4.0f * dot(V, H));
  float G = SmithG_GGX(fabsf(dot(N, L)), m.roughness) * SmithG_GGX(dot(N, V), m.roughness);
  return m.albedo * F * D * G;
}
VKRT_D float3 DisneyDielectricRefraction(const ShadeMat& m, float eta, float3 V, float3 N, float3 L, float3 H, float& pdf)
{
  float F         = DielectricFresnel(fabsf(dot(V, H)), eta);
  float D         = GTR2(dot(N, H), m.roughness);
  float denomSqrt = dot(L, H) * eta + dot(V, H);
  pdf             = D * dot(N, H) * (1.0f - F) * fabsf(dot(L, H)) / (denomSqrt * denomSqrt);
  float G = SmithG_GGX(fabsf(dot(N, L)), m.roughness) * SmithG_GGX(dot(N, V), m.roughness);
  return m.albedo * (1.0f - F) * D * G * fabsf(dot(V, H)) * fabsf(dot(L, H)) * 4.0f * eta * eta / (denomSqrt * denomSqrt);
}
VKRT_D float3 DisneySpecular(const ShadeMat& m, const ShadeFrame& fr, float3 Cspec0, float3 V, float3 N, float3 L, float3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return f3(0.0f);
  float D  = GTR2_aniso(dot(N, H), dot(H, fr.tangent), dot(H, fr.bitangent), m.ax, m.ay);
  pdf      = D * dot(N, H) / (4.0f * dot(V, H));
  float  FH = SchlickFresnel(dot(L, H));
  float3 F  = mix(Cspec0, f3(1.0f), FH);
  float  G  = SmithG_GGX_aniso(dot(N, L), dot(L, fr.tangent), dot(L, fr.bitangent), m.ax, m.ay);
  G *= SmithG_GGX_aniso(dot(N, V), dot(V, fr.tangent), dot(V, fr.bitangent), m.ax, m.ay);
  return F * D * G;
}
VKRT_D float3 DisneyClearcoat(const ShadeMat& m, float3 V, float3 N, float3 L, float3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return f3(0.0f);
  float D  = GTR1(dot(N, H), m.clearcoatRoughness);
  pdf      = D * dot(N, H) / (4.0f * dot(V, H));
  float FH = SchlickFresnel(dot(L, H));
  float F  = mixf(0.04f, 1.0f, FH);
  float G  = SmithG_GGX(dot(N, L), 0.25f) * SmithG_GGX(dot(N, V), 0.25f);
  return f3(0.25f * m.clearcoat * F * D * G);
}
VKRT_D float3 DisneyDiffuse(const ShadeMat& m, float3 Csheen, float3 V, float3 N, float3 L, float3 H, float& pdf)
{
  if(dot(N, L) < 0.0f)
    return f3(0.0f);
  pdf           = dot(N, L) * (1.0f / kPi);
  float  FL     = SchlickFresnel(dot(N, L));
  float  FV     = SchlickFresnel(dot(N, V));
  float  FH     = SchlickFresnel(dot(L, H));
  float  Fd90   = 0.5f + 2.0f * dot(L, H) * dot(L, H) * m.roughness;
  float  Fd     = mixf(1.0f, Fd90, FL) * mixf(1.0f, Fd90, FV);
  float3 Fsheen = FH * m.sheen * Csheen;
  return ((1.0f / kPi) * Fd * (1.0f - m.subsurface) * m.albedo + Fsheen) * (1.0f - m.metallic);
}
VKRT_D float3 DisneySubsurface(const ShadeMat& m, float3 V, float3 N, float3 L, float& pdf)
{
  pdf      = (1.0f / kTwoPi);
  float FL = SchlickFresnel(fabsf(dot(N, L)));
  float FV = SchlickFresnel(dot(N, V));
  float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
  return sqrt3(m.albedo) * m.subsurface * (1.0f / kPi) * Fd * (1.0f - m.metallic) * (1.0f - m.transmission);
}

VKRT_D void DisneyTints(const ShadeMat& m, float3& Cspec0, float3& Csheen)
{
  float3 Cdlin = m.albedo;
  float  Cdlum = 0.3f * Cdlin.x + 0.6f * Cdlin.y + 0.1f * Cdlin.z;
  float3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : f3(1.0f);
  Cspec0       = mix(m.specular * 0.08f * mix(f3(1.0f), Ctint, m.specularTint), Cdlin, m.metallic);
  Csheen       = m.sheenTint;
}

// pbr_disney.glsl:414-520.  `fr` is inout: eta and isSubsurface are written (used by the RR term).
VKRT_D float3 DisneySample(const ShadeMat& m, ShadeFrame& fr, float3 V, float3 N, float3& L, float& pdf, uint32_t& seed)
{
  fr.isSubsurface = false;
  pdf             = 0.0f;
  float3 f        = f3(0.0f);

  float r1 = rnd(seed);
  float r2 = rnd(seed);

  float diffuseRatio = 0.5f * (1.0f - m.metallic);
  float transWeight  = (1.0f - m.metallic) * m.transmission;
  float3 Cspec0, Csheen;
  DisneyTints(m, Cspec0, Csheen);

  if(rnd(seed) < transWeight)
  {
    float3 H = SampleGTR2(m.roughness, r1, r2);
    H        = fr.tangent * H.x + fr.bitangent * H.y + N * H.z;
    float3 R = reflect(-V, H);
    float  F = DielectricFresnel(fabsf(dot(R, H)), fr.eta);
    if(m.thinwalled)
    {
      if(dot(fr.ffnormal, fr.normal) < 0.0f)
        F = 0;
      fr.eta = 1.001f;
    }
    if(rnd(seed) < F)
    {
      L = normalize(R);
      f = DisneyDielectricReflection(m, fr.eta, V, N, L, H, pdf);
    }
    else
    {
      L = normalize(refract(-V, H, fr.eta));
      f = DisneyDielectricRefraction(m, fr.eta, V, N, L, H, pdf);
    }
    f *= transWeight;
    pdf *= transWeight;
  }
  else
  {
    if(rnd(seed) < diffuseRatio)
    {
      if(rnd(seed) < m.subsurface)
      {
        L = UniformSampleHemisphere(r1, r2);
        L = fr.tangent * L.x + fr.bitangent * L.y - N * L.z;
        f = DisneySubsurface(m, V, N, L, pdf);
        pdf *= m.subsurface * diffuseRatio;
        fr.isSubsurface = true;
      }
      else
      {
        L        = CosineSampleHemisphere(r1, r2);
        L        = fr.tangent * L.x + fr.bitangent * L.y + N * L.z;
        float3 H = normalize(L + V);
        f        = DisneyDiffuse(m, Csheen, V, N, L, H, pdf);
        pdf *= (1.0f - m.subsurface) * diffuseRatio;
      }
    }
    else
    {
      float primarySpecRatio = 1.0f / (1.0f + m.clearcoat);
      if(rnd(seed) < primarySpecRatio)
      {
        float3 H = SampleGTR2Aniso(m.ax, m.ay, r1, r2);
        H        = fr.tangent * H.x + fr.bitangent * H.y + N * H.z;
        L        = normalize(reflect(-V, H));
        f        = DisneySpecular(m, fr, Cspec0, V, N, L, H, pdf);
        pdf *= primarySpecRatio * (1.0f - diffuseRatio);
      }
      else
      {
        float3 H = SampleGTR1(m.clearcoatRoughness, r1, r2);
        H        = fr.tangent * H.x + fr.bitangent * H.y + N * H.z;
        L        = normalize(reflect(-V, H));
        f        = DisneyClearcoat(m, V, N, L, H, pdf);
        pdf *= (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
      }
    }
    f *= (1.0f - transWeight);
    pdf *= (1.0f - transWeight);
  }
  return f;
}

// pbr_disney.glsl:524-599
VKRT_D float3 DisneyEval(const ShadeMat& m, const ShadeFrame& fr, float3 V, float3 N, float3 L, float& pdf)
{
  float3 H;
  if(dot(N, L) < 0.0f)
    H = normalize(L * (1.0f / fr.eta) + V);
  else
    H = normalize(L + V);
  if(dot(N, H) < 0.0f)
    H = -H;

  float diffuseRatio     = 0.5f * (1.0f - m.metallic);
  float primarySpecRatio = 1.0f / (1.0f + m.clearcoat);
  float transWeight      = (1.0f - m.metallic) * m.transmission;

  float3 brdf = f3(0.0f), bsdf = f3(0.0f);
  float  brdfPdf = 0.0f, bsdfPdf = 0.0f;

  if(transWeight > 0.0f)
  {
    if(dot(N, L) < 0.0f)
      bsdf = DisneyDielectricRefraction(m, fr.eta, V, N, L, H, bsdfPdf);
    else
      bsdf = DisneyDielectricReflection(m, fr.eta, V, N, L, H, bsdfPdf);
  }
  float m_pdf = 0.0f;
  if(transWeight < 1.0f)
  {
    if(dot(N, L) < 0.0f)
    {
      if(m.subsurface > 0.0f)
      {
        brdf    = DisneySubsurface(m, V, N, L, m_pdf);
        brdfPdf = m_pdf * m.subsurface * diffuseRatio;
      }
    }
    else
    {
      float3 Cspec0, Csheen;
      DisneyTints(m, Cspec0, Csheen);
      brdf += DisneyDiffuse(m, Csheen, V, N, L, H, m_pdf);
      brdfPdf += m_pdf * (1.0f - m.subsurface) * diffuseRatio;
      brdf += DisneySpecular(m, fr, Cspec0, V, N, L, H, m_pdf);
      brdfPdf += m_pdf * primarySpecRatio * (1.0f - diffuseRatio);
      brdf += DisneyClearcoat(m, V, N, L, H, m_pdf);
      brdfPdf += m_pdf * (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
    }
  }
  pdf = mixf(brdfPdf, bsdfPdf, transWeight);
  return mix(brdf, bsdf, transWeight);
}

// ----------------------------------------------------------------------------- glTF model
// pbr_gltf.glsl:36-112
VKRT_D float  pow5(float x) { return m_powf(x, 5.0f); }
VKRT_D float3 F_Schlick3(float3 f0, float3 f90, float VdotH) { return f0 + (f90 - f0) * pow5(clampf(1.0f - VdotH, 0.0f, 1.0f)); }
VKRT_D float  F_Schlick1(float f0, float f90, float VdotH) { return f0 + (f90 - f0) * pow5(clampf(1.0f - VdotH, 0.0f, 1.0f)); }
VKRT_D float  V_GGX(float NdotL, float NdotV, float alphaRoughness)
{
  float a2   = alphaRoughness * alphaRoughness;
  float GGXV = NdotL * sqrtf(NdotV * NdotV * (1.0f - a2) + a2);
  float GGXL = NdotV * sqrtf(NdotL * NdotL * (1.0f - a2) + a2);
  float GGX  = GGXV + GGXL;
  return GGX > 0.0f ? 0.5f / GGX : 0.0f;
}
VKRT_D float V_GGX_anisotropic(float NdotL, float NdotV, float BdotV, float TdotV, float TdotL, float BdotL, float at, float ab)
{
  float GGXV = NdotL * length(f3(at * TdotV, ab * BdotV, NdotV));
  float GGXL = NdotV * length(f3(at * TdotL, ab * BdotL, NdotL));
  float v    = 0.5f / (GGXV + GGXL);
  return clampf(v, 0.0f, 1.0f);
}
VKRT_D float D_GGX(float NdotH, float alphaRoughness)
{
  float a2 = alphaRoughness * alphaRoughness;
  float f  = (NdotH * NdotH) * (a2 - 1.0f) + 1.0f;
  return a2 / (kPi * f * f);
}
VKRT_D float D_GGX_anisotropic(float NdotH, float TdotH, float BdotH, float at, float ab)
{
  float  a2 = at * ab;
  float3 f  = f3(ab * TdotH, at * BdotH, a2 * NdotH);
  float  w2 = a2 / dot(f, f);
  return a2 * w2 * w2 / kPi;
}
// pbr_gltf.glsl:139-146
VKRT_D float3 BRDF_specularGGX(float3 f0, float3 f90, float alphaRoughness, float VdotH, float NdotL, float NdotV, float NdotH)
{
  float3 F = F_Schlick3(f0, f90, VdotH);
  float  Vis = V_GGX(NdotL, NdotV, alphaRoughness);
  float  D = D_GGX(NdotH, fmaxf(0.001f, alphaRoughness));
  return F * Vis * D;
}
// pbr_gltf.glsl:188-198
VKRT_D float3 GgxSampling(float specularAlpha, float r1, float r2)
{
  float phi      = r1 * 2.0f * kPi;
  float cosTheta = sqrtf((1.0f - r2) / (1.0f + (specularAlpha * specularAlpha - 1.0f) * r2));
  float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
  float sinPhi = m_sinf(phi), cosPhi = m_cosf(phi);
  return f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
// pbr_gltf.glsl:203-219
VKRT_D float3 GltfDiffuse(const ShadeMat& m, float3 V, float3 N, float3 L, float& pdf)
{
  pdf         = 0;
  float NdotV = dot(N, V);
  float NdotL = dot(N, L);
  if(NdotL < 0.0f || NdotV < 0.0f)
    return f3(0.0f);
  NdotL = clampf(NdotL, 0.001f, 1.0f);
  pdf   = NdotL * kInvPi;
  return (1.0f - m.metallic) * (m.albedo / kPi);
}
// pbr_gltf.glsl:224-285
VKRT_D float3 GltfSpecular(const ShadeMat& m, const ShadeFrame& fr, float3 f0, float3 f90, float3 V, float3 N, float3 L, float3 H, float& pdf)
{
  pdf         = 0;
  float NdotL = dot(N, L);
  if(NdotL < 0.0f)
    return f3(0.0f);
  if(m.anisotropy > 0)
  {
    float3 T = fr.tangent, B = fr.bitangent;
    float  TdotV = clampf(dot(T, V), 0.f, 1.f);
    float  BdotV = clampf(dot(B, V), 0.f, 1.f);
    float  TdotL = dot(T, L), BdotL = dot(B, L), TdotH = dot(T, H), BdotH = dot(B, H);
    float  NdotH = dot(N, H), NdotV = dot(N, V), VdotH = dot(V, H), LdotH = dot(L, H);
    NdotL = clampf(NdotL, 0.001f, 1.0f);
    NdotV = clampf(fabsf(NdotV), 0.001f, 1.0f);
    float at = fmaxf(m.roughness * (1.0f + m.anisotropy), 0.001f);
    float ab = fmaxf(m.roughness * (1.0f - m.anisotropy), 0.001f);
    pdf      = D_GGX_anisotropic(NdotH, TdotH, BdotH, at, ab) / (4.0f * LdotH);
    // BRDF_specularAnisotropicGGX (pbr_gltf.glsl:149-175)
    float  at2 = fmaxf(m.roughness * (1.0f + m.anisotropy), 0.00001f);
    float  ab2 = fmaxf(m.roughness * (1.0f - m.anisotropy), 0.00001f);
    float3 F   = F_Schlick3(f0, f90, VdotH);
    float  Vis = V_GGX_anisotropic(NdotL, NdotV, BdotV, TdotV, TdotL, BdotL, at2, ab2);
    float  D   = D_GGX_anisotropic(NdotH, TdotH, BdotH, at2, ab2);
    return F * Vis * D;
  }
  float NdotV = dot(N, V);
  float NdotH = clampf(dot(N, H), 0.f, 1.f);
  float LdotH = clampf(dot(L, H), 0.f, 1.f);
  float VdotH = clampf(dot(V, H), 0.f, 1.f);
  NdotL = clampf(NdotL, 0.001f, 1.0f);
  NdotV = clampf(fabsf(NdotV), 0.001f, 1.0f);
  pdf   = D_GGX(NdotH, m.roughness) * NdotH / (4.0f * LdotH);
  return BRDF_specularGGX(f0, f90, m.roughness, VdotH, NdotL, NdotV, NdotH);
}
// pbr_gltf.glsl:290-315
VKRT_D float3 GltfClearcoat(const ShadeMat& m, float3 V, float3 N, float3 L, float3 H, float& pdf)
{
  pdf         = 0;
  float NdotL = dot(N, L);
  if(NdotL < 0.0f)
    return f3(0.0f);
  float NdotH = dot(N, H), NdotV = dot(N, V), VdotH = dot(V, H), LdotH = dot(L, H);
  NdotL = clampf(NdotL, 0.001f, 1.0f);
  NdotV = clampf(fabsf(NdotV), 0.001f, 1.0f);
  float clearcoatFresnel = F_Schlick1(0.04f, 1.f, VdotH);
  float clearcoatAlpha   = m.clearcoatRoughness * m.clearcoatRoughness;
  float G                = V_GGX(NdotL, NdotV, clearcoatAlpha);
  float D                = D_GGX(NdotH, fmaxf(0.001f, clearcoatAlpha));
  pdf                    = D * NdotH / (4.0f * LdotH);
  return f3(clearcoatFresnel * D * G * m.clearcoat);
}
VKRT_D void GltfF0F90(const ShadeMat& m, float3& f0, float3& f90)
{
  float reflectance = fmaxf(fmaxf(m.f0.x, m.f0.y), m.f0.z);
  f0                = m.f0;
  f90               = f3(clampf(reflectance * 50.0f, 0.0f, 1.0f));
}
// pbr_gltf.glsl:365-434
VKRT_D float3 PbrEval(const ShadeMat& m, const ShadeFrame& fr, float3 V, float3 N, float3 L, float& pdfOut)
{
  float3 H;
  if(dot(N, L) < 0.0f)
    H = normalize(L * (1.0f / fr.eta) + V);
  else
    H = normalize(L + V);
  if(dot(N, H) < 0.0f)
    H = -H;
  float  transWeight = (1.0f - m.metallic) * m.transmission;
  float3 brdf = f3(0.0f), bsdf = f3(0.0f);
  float  brdfPdf = 0.0f, bsdfPdf = 0.0f;
  if(transWeight > 0.0f)
  {
    // EvalDielectricRefractionGltf (pbr_gltf.glsl:341-345): pass-through
    bsdfPdf = fabsf(dot(N, L));
    bsdf    = m.albedo;
  }
  if(transWeight < 1.0f && dot(N, L) > 0)
  {
    float  pdf;
    float  diffuseRatio     = 0.5f * (1.0f - m.metallic);
    float  specularRatio    = 1.0f - diffuseRatio;
    float  primarySpecRatio = 1.0f / (1.0f + m.clearcoat);
    float3 f0, f90;
    GltfF0F90(m, f0, f90);
    brdf += GltfDiffuse(m, V, N, L, pdf);
    brdfPdf += pdf * diffuseRatio;
    brdf += GltfClearcoat(m, V, N, L, H, pdf);
    brdfPdf += pdf * (1.0f - primarySpecRatio) * specularRatio;
    brdf += GltfSpecular(m, fr, f0, f90, V, N, L, H, pdf);
    brdfPdf += pdf * primarySpecRatio * specularRatio;
  }
  pdfOut = mixf(brdfPdf, bsdfPdf, transWeight);
  return mix(brdf, bsdf, transWeight);
}
// pbr_gltf.glsl:439-554 (State is `in`: nothing is written back)
VKRT_D float3 PbrSample(const ShadeMat& m, const ShadeFrame& fr, float3 V, float3 N, float3& L, float& pdf, uint32_t& seed)
{
  pdf         = 0.0f;
  float3 brdf = f3(0.0f);

  float probability   = rnd(seed);
  float diffuseRatio  = 0.5f * (1.0f - m.metallic);
  float specularRatio = 1.0f - diffuseRatio;
  float transWeight   = (1.0f - m.metallic) * m.transmission;
  float r1 = rnd(seed);
  float r2 = rnd(seed);

  if(rnd(seed) < transWeight)
  {
    float  eta = fr.eta;
    float  n1 = 1.0f, n2 = m.ior;
    float  R0 = (n1 - n2) / (n1 + n2);
    float3 H  = GgxSampling(m.roughness, r1, r2);
    H         = fr.tangent * H.x + fr.bitangent * H.y + N * H.z;
    float VdotH       = dot(V, H);
    float F           = F_Schlick1(R0 * R0, 1.0f, VdotH);
    float discriminat = 1.0f - eta * eta * (1.0f - VdotH * VdotH);
    if(m.thinwalled)
    {
      if(dot(fr.ffnormal, fr.normal) < 0.0f)
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
      if(isnan(L.x) || isnan(L.y) || isnan(L.z))
        L = -V;
    }
    pdf  = fabsf(dot(N, L));
    brdf = m.albedo;
  }
  else
  {
    float3 f0, f90;
    GltfF0F90(m, f0, f90);
    float3 T = fr.tangent, B = fr.bitangent;
    if(probability < diffuseRatio)
    {
      L    = CosineSampleHemisphere(r1, r2);
      L    = L.x * T + L.y * B + L.z * N;
      brdf = GltfDiffuse(m, V, N, L, pdf);
      pdf *= (1.0f - m.subsurface) * diffuseRatio;
    }
    else
    {
      float primarySpecRatio = 1.0f / (1.0f + m.clearcoat);
      float roughness        = (rnd(seed) < primarySpecRatio) ? m.roughness : m.clearcoatRoughness;
      float3 H = GgxSampling(roughness, r1, r2);
      H        = T * H.x + B * H.y + N * H.z;
      L        = reflect(-V, H);
      if(rnd(seed) < primarySpecRatio)
      {
        brdf = GltfSpecular(m, fr, f0, f90, V, N, L, H, pdf);
        pdf *= primarySpecRatio * specularRatio;
      }
      else
      {
        brdf = GltfClearcoat(m, V, N, L, H, pdf);
        pdf *= (1.0f - primarySpecRatio) * specularRatio;
      }
    }
    brdf *= (1.0f - transWeight);
    pdf *= (1.0f - transWeight);
  }
  return brdf;
}

}  // namespace vkrt
