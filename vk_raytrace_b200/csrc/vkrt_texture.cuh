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
  // power-of-two sizes (the common case) wrap with a mask: the same value as the signed modulo below, without the
  // ~30-instruction integer division (two's complement: (-1) & (n - 1) == n - 1)
  if((n & (n - 1)) == 0)
  {
    if(mode == 2)
    {
      const int m = i & (2 * n - 1);
      return m < n ? m : 2 * n - 1 - m;
    }
    return i & (n - 1);
  }
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

// UNORM8 -> float: k / 255.0f for k = 0..255, folded by the compiler with IEEE round-to-nearest division -- bit-identical to
// dividing on the device, but one cached load instead of a division.  (ncu, round 1: the alpha test spent a quarter of its
// ~440 instructions in the slow path of the division routine, which a zero numerator -- a fully transparent texel -- always takes.)
static __device__ const float kUnorm8[256] = {
  0.0f / 255.0f, 1.0f / 255.0f, 2.0f / 255.0f, 3.0f / 255.0f, 4.0f / 255.0f, 5.0f / 255.0f, 6.0f / 255.0f, 7.0f / 255.0f,
  8.0f / 255.0f, 9.0f / 255.0f, 10.0f / 255.0f, 11.0f / 255.0f, 12.0f / 255.0f, 13.0f / 255.0f, 14.0f / 255.0f, 15.0f / 255.0f,
  16.0f / 255.0f, 17.0f / 255.0f, 18.0f / 255.0f, 19.0f / 255.0f, 20.0f / 255.0f, 21.0f / 255.0f, 22.0f / 255.0f, 23.0f / 255.0f,
  24.0f / 255.0f, 25.0f / 255.0f, 26.0f / 255.0f, 27.0f / 255.0f, 28.0f / 255.0f, 29.0f / 255.0f, 30.0f / 255.0f, 31.0f / 255.0f,
  32.0f / 255.0f, 33.0f / 255.0f, 34.0f / 255.0f, 35.0f / 255.0f, 36.0f / 255.0f, 37.0f / 255.0f, 38.0f / 255.0f, 39.0f / 255.0f,
  40.0f / 255.0f, 41.0f / 255.0f, 42.0f / 255.0f, 43.0f / 255.0f, 44.0f / 255.0f, 45.0f / 255.0f, 46.0f / 255.0f, 47.0f / 255.0f,
  48.0f / 255.0f, 49.0f / 255.0f, 50.0f / 255.0f, 51.0f / 255.0f, 52.0f / 255.0f, 53.0f / 255.0f, 54.0f / 255.0f, 55.0f / 255.0f,
  56.0f / 255.0f, 57.0f / 255.0f, 58.0f / 255.0f, 59.0f / 255.0f, 60.0f / 255.0f, 61.0f / 255.0f, 62.0f / 255.0f, 63.0f / 255.0f,
  64.0f / 255.0f, 65.0f / 255.0f, 66.0f / 255.0f, 67.0f / 255.0f, 68.0f / 255.0f, 69.0f / 255.0f, 70.0f / 255.0f, 71.0f / 255.0f,
  72.0f / 255.0f, 73.0f / 255.0f, 74.0f / 255.0f, 75.0f / 255.0f, 76.0f / 255.0f, 77.0f / 255.0f, 78.0f / 255.0f, 79.0f / 255.0f,
  80.0f / 255.0f, 81.0f / 255.0f, 82.0f / 255.0f, 83.0f / 255.0f, 84.0f / 255.0f, 85.0f / 255.0f, 86.0f / 255.0f, 87.0f / 255.0f,
  88.0f / 255.0f, 89.0f / 255.0f, 90.0f / 255.0f, 91.0f / 255.0f, 92.0f / 255.0f, 93.0f / 255.0f, 94.0f / 255.0f, 95.0f / 255.0f,
  96.0f / 255.0f, 97.0f / 255.0f, 98.0f / 255.0f, 99.0f / 255.0f, 100.0f / 255.0f, 101.0f / 255.0f, 102.0f / 255.0f, 103.0f / 255.0f,
  104.0f / 255.0f, 105.0f / 255.0f, 106.0f / 255.0f, 107.0f / 255.0f, 108.0f / 255.0f, 109.0f / 255.0f, 110.0f / 255.0f, 111.0f / 255.0f,
  112.0f / 255.0f, 113.0f / 255.0f, 114.0f / 255.0f, 115.0f / 255.0f, 116.0f / 255.0f, 117.0f / 255.0f, 118.0f / 255.0f, 119.0f / 255.0f,
  120.0f / 255.0f, 121.0f / 255.0f, 122.0f / 255.0f, 123.0f / 255.0f, 124.0f / 255.0f, 125.0f / 255.0f, 126.0f / 255.0f, 127.0f / 255.0f,
  128.0f / 255.0f, 129.0f / 255.0f, 130.0f / 255.0f, 131.0f / 255.0f, 132.0f / 255.0f, 133.0f / 255.0f, 134.0f / 255.0f, 135.0f / 255.0f,
  136.0f / 255.0f, 137.0f / 255.0f, 138.0f / 255.0f, 139.0f / 255.0f, 140.0f / 255.0f, 141.0f / 255.0f, 142.0f / 255.0f, 143.0f / 255.0f,
  144.0f / 255.0f, 145.0f / 255.0f, 146.0f / 255.0f, 147.0f / 255.0f, 148.0f / 255.0f, 149.0f / 255.0f, 150.0f / 255.0f, 151.0f / 255.0f,
  152.0f / 255.0f, 153.0f / 255.0f, 154.0f / 255.0f, 155.0f / 255.0f, 156.0f / 255.0f, 157.0f / 255.0f, 158.0f / 255.0f, 159.0f / 255.0f,
  160.0f / 255.0f, 161.0f / 255.0f, 162.0f / 255.0f, 163.0f / 255.0f, 164.0f / 255.0f, 165.0f / 255.0f, 166.0f / 255.0f, 167.0f / 255.0f,
  168.0f / 255.0f, 169.0f / 255.0f, 170.0f / 255.0f, 171.0f / 255.0f, 172.0f / 255.0f, 173.0f / 255.0f, 174.0f / 255.0f, 175.0f / 255.0f,
  176.0f / 255.0f, 177.0f / 255.0f, 178.0f / 255.0f, 179.0f / 255.0f, 180.0f / 255.0f, 181.0f / 255.0f, 182.0f / 255.0f, 183.0f / 255.0f,
  184.0f / 255.0f, 185.0f / 255.0f, 186.0f / 255.0f, 187.0f / 255.0f, 188.0f / 255.0f, 189.0f / 255.0f, 190.0f / 255.0f, 191.0f / 255.0f,
  192.0f / 255.0f, 193.0f / 255.0f, 194.0f / 255.0f, 195.0f / 255.0f, 196.0f / 255.0f, 197.0f / 255.0f, 198.0f / 255.0f, 199.0f / 255.0f,
  200.0f / 255.0f, 201.0f / 255.0f, 202.0f / 255.0f, 203.0f / 255.0f, 204.0f / 255.0f, 205.0f / 255.0f, 206.0f / 255.0f, 207.0f / 255.0f,
  208.0f / 255.0f, 209.0f / 255.0f, 210.0f / 255.0f, 211.0f / 255.0f, 212.0f / 255.0f, 213.0f / 255.0f, 214.0f / 255.0f, 215.0f / 255.0f,
  216.0f / 255.0f, 217.0f / 255.0f, 218.0f / 255.0f, 219.0f / 255.0f, 220.0f / 255.0f, 221.0f / 255.0f, 222.0f / 255.0f, 223.0f / 255.0f,
  224.0f / 255.0f, 225.0f / 255.0f, 226.0f / 255.0f, 227.0f / 255.0f, 228.0f / 255.0f, 229.0f / 255.0f, 230.0f / 255.0f, 231.0f / 255.0f,
  232.0f / 255.0f, 233.0f / 255.0f, 234.0f / 255.0f, 235.0f / 255.0f, 236.0f / 255.0f, 237.0f / 255.0f, 238.0f / 255.0f, 239.0f / 255.0f,
  240.0f / 255.0f, 241.0f / 255.0f, 242.0f / 255.0f, 243.0f / 255.0f, 244.0f / 255.0f, 245.0f / 255.0f, 246.0f / 255.0f, 247.0f / 255.0f,
  248.0f / 255.0f, 249.0f / 255.0f, 250.0f / 255.0f, 251.0f / 255.0f, 252.0f / 255.0f, 253.0f / 255.0f, 254.0f / 255.0f, 255.0f / 255.0f,
};
VKRT_D float unorm8(unsigned char k) { return __ldg(&kUnorm8[k]); }

VKRT_D float4 texel8(const TexDesc& t, int x, int y)
{
  uchar4 p = __ldg(&t.data[(size_t)y * t.w + x]);
  return f4(unorm8(p.x), unorm8(p.y), unorm8(p.z), unorm8(p.w));
}
// alpha channel only (the any-hit alpha test)
VKRT_D float texel8Alpha(const TexDesc& t, int x, int y) { return unorm8(__ldg(&t.data[(size_t)y * t.w + x]).w); }

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
