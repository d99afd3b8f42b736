// vkrt_sunsky.cuh -- analytic sun & sky environment for the shade kernel
// (reference shaders/sun_and_sky.glsl:29-601; active when SunAndSky.in_use == 1,
// default off: src/sample_example.hpp:192).  Kept out of line (__noinline__) so the
// common HDR-environment path of the shade kernel does not pay its register cost.
#pragma once
#include "vkrt_math.cuh"
#include "../../include/vkrt_host_device.h"

namespace vkrt {
namespace sky {

VKRT_D float lum(float3 rgb) { return 0.2126f * rgb.x + 0.7152f * rgb.y + 0.0722f * rgb.z; }

// sun_and_sky.glsl:36-70
VKRT_D float3 xyz2dir(float3 n, float x, float y, float z)
{
  float3 u;
  if(fabsf(n.x) < fabsf(n.y))
    u = f3(0.0f, -n.z, n.y);
  else
    u = f3(n.z, 0.0f, -n.x);
  // (the reference's degenerate-transform retry recomputes the same vector)
  u        = normalize(u);
  float3 v = cross(n, u);
  return x * u + y * v + z * n;
}
// sun_and_sky.glsl:73-115
VKRT_D float2 squareToDisk(float inX, float inY)
{
  float r = 0.f, phi = 0.f;
  float lx = 2 * inX - 1, ly = 2 * inY - 1;
  if(!(lx == 0.0f && ly == 0.0f))
  {
    if(lx > -ly)
    {
      if(lx > ly) { r = lx; phi = (kPi / 4.0f) * (1.0f + ly / lx); }
      else { r = ly; phi = (kPi / 4.0f) * (3.0f - lx / ly); }
    }
    else
    {
      if(lx < ly) { r = -lx; phi = (kPi / 4.0f) * (5.0f + ly / lx); }
      else { r = -ly; phi = (kPi / 4.0f) * (7.0f - lx / ly); }
    }
  }
  return f2(r, phi);
}
// sun_and_sky.glsl:118-138
VKRT_D float3 diffuseDir(float3 n, float2 smp)
{
  float2 rp = squareToDisk(smp.x, smp.y);
  float  x  = rp.x * cosf(rp.y);
  float  y  = rp.x * sinf(rp.y);
  float  z2 = 1.0f - x * x - y * y;
  float  z  = z2 > 0.0f ? sqrtf(z2) : 0.0f;
  return xyz2dir(n, x, y, z);
}
// sun_and_sky.glsl:141-165
VKRT_D float3 sunColor(float3 sunDir, float turbidity)
{
  float3 c = f3(0.0f);
  const float3 ko = f3(12.0f, 8.5f, 0.9f), wl = f3(0.610f, 0.550f, 0.470f);
  const float3 solRad = f3(1.0f * 127500 / 0.9878f, 0.992f * 127500 / 0.9878f, 0.911f * 127500 / 0.9878f);
  if(sunDir.z > 0.0f)
  {
    float  m     = (1.0f / (sunDir.z + 0.15f * powf(93.885f - acosf(sunDir.z) * 180 / kPi, -1.253f)));
    float  beta  = 0.04608f * turbidity - 0.04586f;
    float  alpha = 1.3f;
    float3 ta    = exp3(-m * beta * pow3(wl, -alpha));
    float  l     = 0.0035f;
    float3 to    = exp3(-m * ko * l);
    float3 tr    = exp3(-m * 0.008735f * pow3(wl, -4.08f));
    c            = tr * ta * to * solRad;
  }
  return c;
}
// sun_and_sky.glsl:168-226
VKRT_D float3 skyColorXyz(float3 dir, float3 sunPos, float T, float luminance)
{
  float3 xyzc;
  float  A, B, C, D, E;
  float  cosGamma = dot(sunPos, dir);
  if(cosGamma > 1.0f)
    cosGamma = 2.0f - cosGamma;
  float gamma       = acosf(cosGamma);
  float cosTheta    = dir.z;
  float cosThetaSun = sunPos.z;
  float thetaSun    = acosf(cosThetaSun);
  float t2 = T * T, ts2 = thetaSun * thetaSun, ts3 = ts2 * thetaSun;
  float zenith_x = ((+0.001650f * ts3 - 0.003742f * ts2 + 0.002088f * thetaSun + 0) * t2
                    + (-0.029028f * ts3 + 0.063773f * ts2 - 0.032020f * thetaSun + 0.003948f) * T
                    + (+0.116936f * ts3 - 0.211960f * ts2 + 0.060523f * thetaSun + 0.258852f));
  float zenith_y = ((+0.002759f * ts3 - 0.006105f * ts2 + 0.003162f * thetaSun + 0) * t2
                    + (-0.042149f * ts3 + 0.089701f * ts2 - 0.041536f * thetaSun + 0.005158f) * T
                    + (+0.153467f * ts3 - 0.267568f * ts2 + 0.066698f * thetaSun + 0.266881f));
  xyzc.y = luminance;
  A = -0.019257f * T - (0.29f - powf(cosThetaSun, 0.5f) * 0.09f);
  B = -0.066513f * T + 0.000818f;
  C = -0.000417f * T + 0.212479f;
  D = -0.064097f * T - 0.898875f;
  E = -0.003251f * T + 0.045178f;
  float x = (((1.f + A * expf(B / cosTheta)) * (1.f + C * expf(D * gamma) + E * cosGamma * cosGamma))
             / ((1 + A * expf(B / 1.0f)) * (1 + C * expf(D * thetaSun) + E * cosThetaSun * cosThetaSun)));
  A = -0.016698f * T - 0.260787f;
  B = -0.094958f * T + 0.009213f;
  C = -0.007928f * T + 0.210230f;
  D = -0.044050f * T - 1.653694f;
  E = -0.010922f * T + 0.052919f;
  float y = (((1 + A * expf(B / cosTheta)) * (1 + C * expf(D * gamma) + E * cosGamma * cosGamma))
             / ((1 + A * expf(B / 1.0f)) * (1 + C * expf(D * thetaSun) + E * cosThetaSun * cosThetaSun)));
  float localSat = 1.0f;
  x = zenith_x * ((x * localSat) + (1.0f - localSat));
  y = zenith_y * ((y * localSat) + (1.0f - localSat));
  xyzc.x = (x / y) * xyzc.y;
  xyzc.z = ((1.0f - x - y) / y) * xyzc.y;
  return xyzc;
}
// sun_and_sky.glsl:229-255
VKRT_D float skyLuminance(float3 dir, float3 sunPos, float T)
{
  float cosGamma = dot(sunPos, dir);
  if(cosGamma < 0.0f)
    cosGamma = 0.0f;
  if(cosGamma > 1.0f)
    cosGamma = 2.0f - cosGamma;
  float gamma       = acosf(cosGamma);
  float cosTheta    = dir.z;
  float cosThetaSun = sunPos.z;
  float thetaSun    = acosf(cosThetaSun);
  float A = 0.178721f * T - 1.463037f;
  float B = -0.355402f * T + 0.427494f;
  float C = -0.022669f * T + 5.325056f;
  float D = 0.120647f * T - 2.577052f;
  float E = -0.066967f * T + 0.370275f;
  return (((1 + A * expf(B / cosTheta)) * (1 + C * expf(D * gamma) + E * cosGamma * cosGamma))
          / ((1 + A * expf(B / 1.0f)) * (1 + C * expf(D * thetaSun) + E * cosThetaSun * cosThetaSun)));
}
// sun_and_sky.glsl:258-273
VKRT_D float3 envColor(float3 sunDir, float3 dir, float T)
{
  float thetaSun = acosf(sunDir.z);
  float chi      = (4.0f / 9.0f - T / 120.0f) * (kPi - 2 * thetaSun);
  float l        = 1000.0f * ((4.0453f * T - 4.9710f) * tanf(chi) - 0.2155f * T + 2.4192f);
  l *= skyLuminance(dir, sunDir, T);
  float3 XYZ = skyColorXyz(dir, sunDir, T, l);
  float3 c   = f3(3.241f * XYZ.x - 1.537f * XYZ.y - 0.499f * XYZ.z, -0.969f * XYZ.x + 1.876f * XYZ.y + 0.042f * XYZ.z,
                  0.056f * XYZ.x - 0.204f * XYZ.y + 1.057f * XYZ.z);
  c *= kPi;
  return c;
}
// sun_and_sky.glsl:275-296
VKRT_D float3 irradiance(float3 sunDir, float haze)
{
  float3       acc = f3(0.0f);
  const float3 n   = f3(0.0f, 0.0f, 1.0f);
  for(float u = 1.f / 10.f; u < 1.f; u += 1.f / 5.f)
    for(float v = 1.f / 10.f; v < 1.f; v += 1.f / 5.f)
      acc += envColor(sunDir, diffuseDir(n, f2(u, v)), haze);
  acc /= 25.0f;
  return acc;
}
// sun_and_sky.glsl:299-316
VKRT_D float tweakSaturation(float sat, float haze)
{
  float lowsat = powf(sat, 3.0f);
  if(sat <= 1.0f)
  {
    float h = haze;
    h -= 2.0f;
    h /= 15.0f;
    if(h < 0.0f) h = 0.0f;
    if(h > 1.0f) h = 1.0f;
    h = powf(h, 3.0f);
    return ((sat * (1.0f - h)) + lowsat * h);
  }
  return 1.f;
}
// sun_and_sky.glsl:319-332
VKRT_D float3 vectorTweak(float3 dir, int yIsUp, float horizHeight)
{
  float3 o = dir;
  if(yIsUp == 1)
    o = f3(dir.x, dir.z, dir.y);
  if(horizHeight != 0)
  {
    o.z -= horizHeight;
    o = normalize(o);
  }
  return o;
}
// sun_and_sky.glsl:335-366
VKRT_D float3 colorTweak(float3 tint, float saturation, float redness)
{
  float  intensity = lum(tint);
  float3 o         = saturation <= 0.0f ? f3(intensity) : tint * saturation + intensity * (1.0f - saturation);
  o *= f3(1.0f + redness, 1.f, 1.0f - redness);
  return o;
}
// sun_and_sky.glsl:369-438
VKRT_D float2 physicalScale(float diskScale, float glowIntensity, float diskIntensity)
{
  float sunAngularRadius = 0.00465f;
  float diskRadius       = sunAngularRadius * diskScale;
  float glowRadius       = diskRadius * 10.0f;
  float glowIntegral     = glowIntensity
                       * ((4.f * kPi) - (24.f * kPi) / (glowRadius * glowRadius)
                          + (24.f * kPi) * sinf(glowRadius) / (glowRadius * glowRadius * glowRadius));
  float targetIntegral = diskIntensity * kPi;
  float glowScale      = 1.0f;
  float maxGlow        = 0.5f * targetIntegral;
  if(glowIntegral > maxGlow)
  {
    glowScale *= maxGlow / glowIntegral;
    targetIntegral -= maxGlow;
  }
  else
  {
    targetIntegral -= glowIntegral;
  }
  float area            = 2 * kPi * (1 - cosf(diskRadius));
  float targetIntensity = targetIntegral / area;
  float actualIntegral  = 1.0f * area;
  float actualIntensity = diskIntensity * 100.0f * actualIntegral / area;
  return f2((targetIntensity == 0.0f) ? 0.0f : targetIntensity / actualIntensity, glowScale);
}
// sun_and_sky.glsl:441-450
VKRT_D float nightBrightness(float3 sunDir)
{
  float lmt = 0.30901699437494742410229341718282f;
  if(sunDir.z <= -lmt)
    return 0.0f;
  float f = (sunDir.z + lmt) / lmt;
  f *= f;
  f *= f;
  return f;
}

// sun_and_sky.glsl:453-601
static __device__ __noinline__ float3 sunAndSky(const SunAndSky& ss, float3 inDirection)
{
  float  factor = 1.0f, nightFactor = 1.0f;
  float3 rgbScale    = f3(ss.rgb_unit_conversion[0], ss.rgb_unit_conversion[1], ss.rgb_unit_conversion[2]);
  float  horizHeight = ss.horizon_height / 10.0f;
  float3 dir         = vectorTweak(inDirection, ss.y_is_up, horizHeight);
  float  haze        = 2.0f + ss.haze;
  if(haze < 2.0f)
    haze = 2.0f;
  float saturation = tweakSaturation(ss.saturation, haze);
  if(lum(rgbScale) < 0.0f)
    rgbScale = f3(1.0f / 80000.0f);
  rgbScale *= ss.multiplier;
  if(ss.multiplier <= 0.0f)
    return f3(0.f);

  float  downness = dir.z;
  float3 realDir  = dir;
  if(dir.z < 0.001f)
  {
    dir.z = 0.001f;
    dir   = normalize(dir);
  }
  float3 sunDir = normalize(f3(ss.sun_direction[0], ss.sun_direction[1], ss.sun_direction[2]));
  sunDir        = vectorTweak(sunDir, ss.y_is_up, horizHeight);
  float3 realSunDir = sunDir;
  if(sunDir.z < 0.001f)
  {
    if(sunDir.z < 0.0f)
      factor = nightBrightness(sunDir);
    sunDir.z = 0.001f;
    sunDir   = normalize(sunDir);
  }
  float3 tint;
  if(factor > 0.0f)
  {
    tint = envColor(sunDir, dir, haze);
    if(factor < 1.0f)
      tint *= factor;
  }
  else
    tint = f3(0.f);
  float3 dataSunColor = sunColor(sunDir, downness > 0 ? haze : 2.0f);
  if(ss.sun_disk_intensity > 0.0f && ss.sun_disk_scale > 0.0f)
  {
    float sunAngle  = acosf(dot(realDir, realSunDir));
    float sunRadius = 0.00465f * ss.sun_disk_scale * 10.0f;
    if(sunAngle < sunRadius)
    {
      float diskScale = 1.0f, glowScale = 1.0f;
      if(ss.physically_scaled_sun == 1)
      {
        float2 rv = physicalScale(ss.sun_disk_scale, ss.sun_glow_intensity, ss.sun_disk_intensity);
        diskScale = rv.x;
        glowScale = rv.y;
      }
      float sunFactor = (1.0f - sunAngle / sunRadius) * 10.0f;
      sunFactor       = (powf(sunFactor / 10.0f, 3.0f) * 2.0f * ss.sun_glow_intensity * glowScale
                   + smoothstepf(8.5f, 9.5f + (haze / 50.0f), sunFactor) * 100.0f * ss.sun_disk_intensity * diskScale);
      tint += dataSunColor * sunFactor;
    }
  }
  float3 outColor = tint * rgbScale;
  if(downness <= 0.0f)
  {
    float3 down  = f3(ss.ground_color[0], ss.ground_color[1], ss.ground_color[2]);
    float3 irrad = irradiance(sunDir, 2.0f);
    down *= (irrad + dataSunColor * sunDir.z) * rgbScale;
    if(factor < 1)
      down *= factor;
    float horBlur = ss.horizon_blur / 10.0f;
    if(horBlur > 0.0f)
    {
      float dness = -downness;
      dness /= horBlur;
      if(dness > 1.0f)
        dness = 1.0f;
      dness       = smoothstepf(0.0f, 1.0f, dness);
      outColor    = outColor * (1.0f - dness) + down * dness;
      nightFactor = 1.0f - dness;
    }
    else
    {
      outColor    = down;
      nightFactor = 0.0f;
    }
  }
  float3 result = colorTweak(outColor, saturation, ss.redblueshift);
  if(nightFactor > 0.0f)
  {
    float3 night = f3(ss.night_color[0], ss.night_color[1], ss.night_color[2]) * nightFactor;
    if(result.x < night.x) result.x = night.x;
    if(result.y < night.y) result.y = night.y;
    if(result.z < night.z) result.z = night.z;
  }
  result *= kPi;
  return result;
}

}  // namespace sky
}  // namespace vkrt
