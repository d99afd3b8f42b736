// vkrt_texture.cuh -- software texture sampling (RGBA8 material textures, RGBA32F
// environment map) at LOD 0, as the reference's textureLod(...,0)/texture() calls in
// compute/ray stages resolve (shaders/gltf_material.glsl:72-183, traceray_rq.glsl:83,
// pathtrace.glsl:224).  Sampler state follows Scene::createTextureImages /
// gltfSamplerToVulkan (src/scene.cpp:447-580) and HdrSampling::loadEnvironment
// (src/hdr_sampling.cpp:68-75: LINEAR, U repeat, V clamp).  Bilinear weights are full
// fp32 (not the 8-bit fixed-point weights of the texture units) so that the result is
// a pure function of the inputs and can be checked exactly against the oracle.
#pragma once
#include "vkrt_device.cuh"

namespace vkrt {

VKRT_D int wrapCoord(int i, int n, int mode)
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

VKRT_D float sanitizeUv(float x) { return (!(x == x) || fabsf(x) > 1e9f) ? 0.0f : x; }

VKRT_D float4 texel8(const TexDesc& t, int x, int y)
{
  uchar4 p = __ldg(&t.data[(size_t)y * t.w + x]);
  return f4(p.x / 255.0f, p.y / 255.0f, p.z / 255.0f, p.w / 255.0f);
}

#ifdef VKRT_COMPACT_TEXTURE
static __device__ __noinline__ float4 sampleTexture(const TexDesc& t, float2 uv)
#else
VKRT_D float4 sampleTexture(const TexDesc& t, float2 uv)
#endif
{
  float u = sanitizeUv(uv.x), v = sanitizeUv(uv.y);
  if(!t.linear)
  {
    int x = wrapCoord((int)floorf(u * (float)t.w), t.w, t.wrapS);
    int y = wrapCoord((int)floorf(v * (float)t.h), t.h, t.wrapT);
    return texel8(t, x, y);
  }
  float fx = u * (float)t.w - 0.5f, fy = v * (float)t.h - 0.5f;
  float x0f = floorf(fx), y0f = floorf(fy);
  float ax = fx - x0f, ay = fy - y0f;
  int   x0 = wrapCoord((int)x0f, t.w, t.wrapS), x1 = wrapCoord((int)x0f + 1, t.w, t.wrapS);
  int   y0 = wrapCoord((int)y0f, t.h, t.wrapT), y1 = wrapCoord((int)y0f + 1, t.h, t.wrapT);
  float4 c00 = texel8(t, x0, y0), c10 = texel8(t, x1, y0), c01 = texel8(t, x0, y1), c11 = texel8(t, x1, y1);
  float4 top = c00 * (1.0f - ax) + c10 * ax;
  float4 bot = c01 * (1.0f - ax) + c11 * ax;
  return top * (1.0f - ay) + bot * ay;
}

#ifdef VKRT_COMPACT_TEXTURE
static __device__ __noinline__ float3 sampleEnv(const SceneView& s, float2 uv)
#else
VKRT_D float3 sampleEnv(const SceneView& s, float2 uv)
#endif
{
  float u = sanitizeUv(uv.x), v = sanitizeUv(uv.y);
  int   w = s.envW, h = s.envH;
  float fx = u * (float)w - 0.5f, fy = v * (float)h - 0.5f;
  float x0f = floorf(fx), y0f = floorf(fy);
  float ax = fx - x0f, ay = fy - y0f;
  int   x0 = wrapCoord((int)x0f, w, 0), x1 = wrapCoord((int)x0f + 1, w, 0);
  int   y0 = wrapCoord((int)y0f, h, 1), y1 = wrapCoord((int)y0f + 1, h, 1);
  float3 c00 = f3(__ldg(&s.envPixels[(size_t)y0 * w + x0])), c10 = f3(__ldg(&s.envPixels[(size_t)y0 * w + x1]));
  float3 c01 = f3(__ldg(&s.envPixels[(size_t)y1 * w + x0])), c11 = f3(__ldg(&s.envPixels[(size_t)y1 * w + x1]));
  float3 top = c00 * (1.0f - ax) + c10 * ax;
  float3 bot = c01 * (1.0f - ax) + c11 * ax;
  return top * (1.0f - ay) + bot * ay;
}

// (vec4(uv,1,1) * uvTransform).xy -- row vector times the column-major mat4
// (shaders/gltf_material.glsl:115, traceray_rq.glsl:81).
VKRT_D float2 applyUvTransform(const float* M, float2 tc)
{
  return f2(((tc.x * M[0] + tc.y * M[1]) + M[2]) + M[3], ((tc.x * M[4] + tc.y * M[5]) + M[6]) + M[7]);
}

}  // namespace vkrt
