// vkrt_shade.cu -- wavefront kernels around the traversal: ray generation, the
// shade / next-bounce kernel, sample finish, accumulation and the tonemapper.
//
// Together they replace the reference's per-pixel megakernel
//   shaders/pathtrace.comp:87-134 / pathtrace.rgen:60-108   (seed, sample loop, accumulate)
//   shaders/pathtrace.glsl:97-387                           (DirectLight, PathTrace, samplePixel)
//   shaders/shade_state.glsl:63-145, gltf_material.glsl:52-193, env_sampling.glsl:38-135,
//   shaders/punctual.glsl:28-51, post.frag:98-147 + tonemapping.glsl:24-65
// The megakernel's `for depth` loop becomes one trace + one shade launch per bounce;
// every path carries its RNG seed and consumes draws in the reference's order
// (SURVEY.md App. B), so results do not depend on scheduling.
#define VKRT_COMPACT_MATH 1
#define VKRT_COMPACT_TEXTURE 1
#include "vkrt_device.cuh"
#include "vkrt_texture.cuh"
#include "vkrt_bsdf.cuh"
#include "vkrt_sunsky.cuh"
#include "vkrt_traverse.cuh"
#include "vkrt_internal.h"

namespace vkrt {

#ifndef VKRT_SHADE_BLOCK
#define VKRT_SHADE_BLOCK 384
#define VKRT_SHADE_MINBLOCKS 2
#endif
constexpr int kShadeBlock     = VKRT_SHADE_BLOCK;
constexpr int kShadeMinBlocks = VKRT_SHADE_MINBLOCKS;  // <= 96 registers: 640 threads per SM; large CTAs so that the phase barriers keep many warps in step

// This file is compiled twice (Makefile): IEEE-exact (everything) and with VKRT_FAST_MATH (the shade kernel only, approximate
// division / square root / SFU transcendentals).  Kernels that exist once live behind !VKRT_FAST_MATH.
#ifndef VKRT_FAST_MATH
// ---- ray generation: pathtrace.comp:97 / pathtrace.rgen:72 + pathtrace.glsl:348-376 -------------------
__global__ void __launch_bounds__(256) raygenKernel(const __grid_constant__ SceneView sc, const __grid_constant__ Wavefront wf,
                                                    const __grid_constant__ FrameParams fp)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p < 4u * wf.clevels)
    wf.counters[p] = (p == 0) ? wf.numPaths : 0u;  // every bounce's counts and cursors (the wavefront always has >= 4 x levels paths or the loop below covers it)
  if(p == 0)
    for(uint32_t k = wf.numPaths; k < 4u * wf.clevels; k++)  // tiny wavefronts: fewer paths than counters
      wf.counters[k] = 0u;
  if(p >= wf.numPaths)
    return;
  const uint32_t f     = p / wf.numLocalPixels;
  const uint32_t lp    = p - f * wf.numLocalPixels;
  const uint32_t pixel = wf.pixelOfLocal[lp];
  const int      W = fp.state.size[0], H = fp.state.size[1];
  const int      px = (int)(pixel % (uint32_t)W), py = (int)(pixel / (uint32_t)W);
  const int      frame = fp.state.frame + (int)f;

  uint32_t seed;
  if(fp.sampleIndex == 0)
  {
    seed = (fp.seedMode == VKRT_SEED_RAYQUERY) ? tea((uint32_t)(W * py + px), (uint32_t)(frame * fp.state.maxSamples))
                                               : tea((uint32_t)py * (uint32_t)W + (uint32_t)px, (uint32_t)frame);
    wf.sum[p] = f4(0.f, 0.f, 0.f, 0.f);
  }
  else
  {
    seed = __float_as_uint(wf.thr[p].w);  // the stream continues across the maxSamples loop
  }

  float2 jitter;
  if(frame == 0)
    jitter = f2(0.5f, 0.5f);
  else
  {
    float jx = rnd(seed);
    float jy = rnd(seed);
    jitter   = f2(jx, jy);
  }
  const float2 pixelCenter = f2((float)px + jitter.x, (float)py + jitter.y);
  const float2 inUV        = f2(pixelCenter.x / (float)W, pixelCenter.y / (float)H);
  const float2 d           = f2(inUV.x * 2.0f - 1.0f, inUV.y * 2.0f - 1.0f);

  const SceneCamera& cam = sc.camera;
  float4 origin    = mulMV(cam.viewInverse, f4(0, 0, 0, 1));
  float4 target    = mulMV(cam.projInverse, f4(d.x, d.y, 1, 1));
  float4 direction = mulMV(cam.viewInverse, f4(normalize(f3(target)), 0));

  float3 focalPoint  = cam.focalDist * f3(direction);
  float  cam_r1      = rnd(seed) * kTwoPi;
  float  cam_r2      = rnd(seed) * cam.aperture;
  float4 cam_right   = mulMV(cam.viewInverse, f4(1, 0, 0, 0));
  float4 cam_up      = mulMV(cam.viewInverse, f4(0, 1, 0, 0));
  float3 aperturePos = (m_cosf(cam_r1) * f3(cam_right) + m_sinf(cam_r1) * f3(cam_up)) * sqrtf(cam_r2);
  float3 rayDir      = normalize(focalPoint - aperturePos);

  stS(&wf.rayO[p], f4(f3(origin) + aperturePos, kInfinity));
  stS(&wf.rayD[p], f4(rayDir, __uint_as_float(seed)));
  stS(&wf.thr[p], f4(1.f, 1.f, 1.f, __uint_as_float(seed)));
  stS(&wf.rad[p], f4(0.f, 0.f, 0.f, 0.f));
  stS(&wf.absb[p], f4(0.f, 0.f, 0.f, 0.f));
  stS(&wf.queueCur[p], p);
}

#endif  // !VKRT_FAST_MATH

// ---- shading state from the hit: shade_state.glsl:63-145 -----------------------------------------------
struct HitGeom
{
  float3   position, normal, tangent, bitangent, color;
  float2   texCoord;
  uint32_t matIndex;
};

VKRT_D HitGeom loadHitGeom(const SceneView& sc, uint32_t triId, float bu, float bv)
{
  HitGeom g;
  const uint4* ts = sc.triShade + 7ull * triId;
  const uint4  a0 = __ldg(ts + 0), a1 = __ldg(ts + 1), b0 = __ldg(ts + 2), b1 = __ldg(ts + 3), c0 = __ldg(ts + 4), c1 = __ldg(ts + 5);
  const uint4  meta = __ldg(ts + 6);  // {instance, material, flags, prim}
  const InstanceXform& I = sc.instances[meta.x];
  // VertexAttributes: {pos.xyz, normal | uv.xy, tangent, color}
  const float3 bary = f3(1.0f - bu - bv, bu, bv);
  const float3 pos0 = f3(__uint_as_float(a0.x), __uint_as_float(a0.y), __uint_as_float(a0.z));
  const float3 pos1 = f3(__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z));
  const float3 pos2 = f3(__uint_as_float(c0.x), __uint_as_float(c0.y), __uint_as_float(c0.z));
  const float3 position = pos0 * bary.x + pos1 * bary.y + pos2 * bary.z;
  g.position            = xformPoint(I.o2w, position);

  float3 normal       = normalize(decompress_unit_vec(a0.w) * bary.x + decompress_unit_vec(b0.w) * bary.y + decompress_unit_vec(c0.w) * bary.z);
  float3 worldNormal  = normalize(xformNormalT(I.w2o, normal));
  float3 geomNormal   = normalize(cross(pos1 - pos0, pos2 - pos0));
  float3 wgeomNormal  = normalize(xformNormalT(I.w2o, geomNormal));

  const float h0 = (a1.y & 1u) ? 1.0f : -1.0f;  // handedness of vertex 0 only (shade_state.glsl:114)
  float3 tangent = decompress_unit_vec(a1.z) * bary.x + decompress_unit_vec(b1.z) * bary.y + decompress_unit_vec(c1.z) * bary.z;
  tangent        = normalize(tangent);
  float3 worldTangent = normalize(xformVector(I.o2w, tangent));
  worldTangent        = normalize(worldTangent - dot(worldTangent, worldNormal) * worldNormal);
  g.bitangent         = cross(worldNormal, worldTangent) * h0;
  g.tangent           = worldTangent;

  const float2 uv0 = f2(__uint_as_float(a1.x), __uint_as_float(a1.y & ~1u));
  const float2 uv1 = f2(__uint_as_float(b1.x), __uint_as_float(b1.y & ~1u));
  const float2 uv2 = f2(__uint_as_float(c1.x), __uint_as_float(c1.y & ~1u));
  g.texCoord       = uv0 * bary.x + uv1 * bary.y + uv2 * bary.z;

  const float4 col = unpackUnorm4x8(a1.w) * bary.x + unpackUnorm4x8(b1.w) * bary.y + unpackUnorm4x8(c1.w) * bary.z;
  g.color          = f3(col);
  g.matIndex       = meta.y;
  g.normal         = worldNormal;
  if(dot(g.normal, wgeomNormal) <= 0)
    g.normal *= -1.0f;
  return g;
}

// ---- material resolve: gltf_material.glsl:52-193 -----------------------------------------------------
VKRT_D float4 srgbToLinear(float4 c) { return f4(pow3(f3(c), 2.2f), c.w); }
VKRT_D float4 tex(const SceneView& sc, int id, float2 uv) { return sampleTexture(sc.textures[id], uv); }

VKRT_D void resolveMaterial(const SceneView& sc, const HitGeom& g, float3 rayDir, ShadeMat& m, ShadeFrame& fr, float2& texCoord)
{
  const GltfShadeMaterial& mat = sc.materials[g.matIndex];
  m.specular     = 0.5f;
  m.subsurface   = 0.f;
  m.specularTint = 1.f;

  texCoord = applyUvTransform(mat.uvTransform, g.texCoord);
  const float3 T0 = g.tangent, B0 = g.bitangent, N0 = g.normal;  // mat3 TBN before normal mapping
  fr.normal    = g.normal;
  fr.tangent   = g.tangent;
  fr.bitangent = g.bitangent;
  fr.ffnormal  = dot(fr.normal, rayDir) <= 0.0f ? fr.normal : -fr.normal;
  fr.isSubsurface = false;

  if(mat.normalTexture > -1)
  {
    float3 nv   = f3(tex(sc, mat.normalTexture, texCoord));
    nv          = normalize(nv * 2.0f - 1.0f);
    nv          = nv * f3(mat.normalTextureScale, mat.normalTextureScale, 1.0f);
    fr.normal   = normalize(T0 * nv.x + B0 * nv.y + N0 * nv.z);
    fr.ffnormal = dot(fr.normal, rayDir) <= 0.0f ? fr.normal : -fr.normal;
    CreateCoordinateSystem(fr.ffnormal, fr.tangent, fr.bitangent);
  }

  m.emission = f3(mat.emissiveFactor[0], mat.emissiveFactor[1], mat.emissiveFactor[2]);
  if(mat.emissiveTexture > -1)
    m.emission *= f3(srgbToLinear(tex(sc, mat.emissiveTexture, texCoord)));

  // GetMetallicRoughness
  float dielectricSpecular = (mat.ior - 1) / (mat.ior + 1);
  dielectricSpecular *= dielectricSpecular;
  float perceptualRoughness = mat.pbrRoughnessFactor;
  float metallic            = mat.pbrMetallicFactor;
  if(mat.pbrMetallicRoughnessTexture > -1)
  {
    float4 mr           = tex(sc, mat.pbrMetallicRoughnessTexture, texCoord);
    perceptualRoughness = mr.y * perceptualRoughness;
    metallic            = mr.z * metallic;
  }
  float4 baseColor = f4(mat.pbrBaseColorFactor[0], mat.pbrBaseColorFactor[1], mat.pbrBaseColorFactor[2], mat.pbrBaseColorFactor[3]);
  if(mat.pbrBaseColorTexture > -1)
    baseColor = baseColor * srgbToLinear(tex(sc, mat.pbrBaseColorTexture, texCoord));
  m.f0        = mix(f3(dielectricSpecular), f3(baseColor), metallic);
  m.albedo    = f3(baseColor);
  m.metallic  = metallic;
  m.alpha     = baseColor.w;
  m.roughness = fmaxf(perceptualRoughness, 0.001f);

  m.transmission = mat.transmissionFactor;
  if(mat.transmissionTexture > -1)
    m.transmission *= tex(sc, mat.transmissionTexture, texCoord).x;

  m.ior  = mat.ior;
  fr.eta = dot(fr.normal, fr.ffnormal) > 0.0f ? (1.0f / m.ior) : m.ior;
  m.unlit = (mat.unlit == 1);

  m.anisotropy = mat.anisotropy;
  float aspect = sqrtf(1.0f - mat.anisotropy * 0.9f);
  m.ax         = fmaxf(0.001f, m.roughness / aspect);
  m.ay         = fmaxf(0.001f, m.roughness * aspect);
  if(mat.anisotropy > 0)
  {
    const float3 ad = f3(mat.anisotropyDirection[0], mat.anisotropyDirection[1], mat.anisotropyDirection[2]);
    fr.tangent      = normalize(T0 * ad.x + B0 * ad.y + N0 * ad.z);
    fr.bitangent    = normalize(cross(fr.normal, fr.tangent));
  }

  m.attenuationColor    = f3(mat.attenuationColor[0], mat.attenuationColor[1], mat.attenuationColor[2]);
  m.attenuationDistance = mat.attenuationDistance;
  m.thinwalled          = mat.thicknessFactor == 0;

  m.clearcoat          = mat.clearcoatFactor;
  m.clearcoatRoughness = mat.clearcoatRoughness;
  if(mat.clearcoatTexture > -1)
    m.clearcoat *= tex(sc, mat.clearcoatTexture, texCoord).x;
  if(mat.clearcoatRoughnessTexture > -1)
    m.clearcoatRoughness *= tex(sc, mat.clearcoatRoughnessTexture, texCoord).y;
  m.clearcoatRoughness = fmaxf(m.clearcoatRoughness, 0.001f);

  float4 sheen = unpackUnorm4x8(mat.sheen);
  m.sheenTint  = f3(sheen);
  m.sheen      = sheen.w;
}

// ---- punctual.glsl:28-51 -----------------------------------------------------------------------------
VKRT_D float rangeAttenuation(float range, float distance)
{
  if(range <= 0.0f)
    return 1.0f;
  return fmaxf(fminf(1.0f - m_powf(distance / range, 4.0f), 1.0f), 0.0f) / m_powf(distance, 2.0f);
}
VKRT_D float spotAttenuation(float3 pointToLight, float3 spotDirection, float outerConeCos, float innerConeCos)
{
  float actualCos = dot(normalize(spotDirection), normalize(-pointToLight));
  if(actualCos > outerConeCos)
  {
    if(actualCos < innerConeCos)
      return smoothstepf(outerConeCos, innerConeCos, actualCos);
    return 1.0f;
  }
  return 0.0f;
}

// ---- env_sampling.glsl:38-135 -------------------------------------------------------------------------
static __device__ __noinline__ float3 environmentSample(const SceneView& sc, float3 xi, float3& toLight, float& pdf)
{
  const uint32_t width = (uint32_t)sc.envW, height = (uint32_t)sc.envH;
  const uint32_t size  = width * height;
  const uint32_t idx   = min((uint32_t)(xi.x * (float)size), size - 1);
  const EnvAccel sd    = sc.envAccel[idx];
  uint32_t       envIdx;
  if(xi.y < sd.q)
  {
    envIdx = idx;
    xi.y /= sd.q;
    pdf = sd.pdf;
  }
  else
  {
    envIdx = sd.alias;
    xi.y   = (xi.y - sd.q) / (1.0f - sd.q);
    pdf    = sd.aliasPdf;
  }
  const uint32_t px = envIdx % width;
  const uint32_t py = envIdx / width;
  const float    u        = (float)((float)px + xi.y) / (float)width;
  const float    phi      = u * (2.0f * kPi) - kPi;
  const float    sinPhi   = m_sinf(phi);
  const float    cosPhi   = m_cosf(phi);
  const float    stepTheta = kPi / (float)height;
  const float    theta0    = (float)py * stepTheta;
  const float    cosTheta  = m_cosf(theta0) * (1.0f - xi.z) + m_cosf(theta0 + stepTheta) * xi.z;
  const float    theta     = m_acosf(cosTheta);
  const float    sinTheta  = m_sinf(theta);
  const float    v         = theta * kInvPi;
  toLight = f3(cosPhi * sinTheta, cosTheta, sinPhi * sinTheta);
  return sampleEnv(sc, f2(u, v));
}

struct ShadeOut
{
  bool   wantNext, wantShadow;
  float3 nextO, nextD;
  float3 shO, shD, shC;
  float  shTmax;
  uint32_t shSeed;
};

// One bounce of pathtrace.glsl:199-340 for one path, cut into three phases so that the block can re-synchronise
// between them: the bounce is ~9k instructions, far beyond the instruction cache, and warps that drift apart
// each stream their own copy of the code from L2 (ncu: "no instruction" was the top stall).  With a barrier
// between phases all warps of a block fetch the same region together.
struct BounceState
{
  float3     rayD, position, radiance, throughput, absorption;
  float      hitT;
  uint32_t   seed;
  ShadeMat   m;
  ShadeFrame fr;
  float3     lightDir, vcontrib;
  float      lightDist;
  bool       visible;
};

// Phase A: miss handling, shading state, material, debug views, emission/absorption (pathtrace.glsl:204-274).
// Returns false when the path ends here.
VKRT_D bool bouncePhaseSurface(const SceneView& sc, const FrameParams& fp, float4 hit, BounceState& b)
{
  const int dbg      = fp.state.debugging_mode;
  const int depth    = fp.depth;
  const int maxDepth = fp.state.maxDepth;
  if(hit.x == kInfinity)
  {
    if(dbg != VKRT_DBG_NONE)
    {
      if(depth != maxDepth - 1) { b.radiance = f3(0.f); return false; }
      if(dbg == VKRT_DBG_RADIANCE) return false;
      else if(dbg == VKRT_DBG_WEIGHT) { b.radiance = b.throughput; return false; }
      else if(dbg == VKRT_DBG_RAYDIR) { b.radiance = (b.rayD + f3(1.f)) * 0.5f; return false; }
    }
    float3 env;
    if(sc.sunsky.in_use == 1)
      env = sky::sunAndSky(sc.sunsky, b.rayD);
    else
      env = sampleEnv(sc, GetSphericalUv(b.rayD));
    b.radiance = b.radiance + (env * fp.state.hdrMultiplier * b.throughput);
    return false;
  }
  const HitGeom g = loadHitGeom(sc, __float_as_uint(hit.w), hit.y, hit.z);
  float2        texCoord;
  resolveMaterial(sc, g, b.rayD, b.m, b.fr, texCoord);
  b.m.albedo *= g.color;
  b.position = g.position;
  b.hitT     = hit.x;
  if(dbg != VKRT_DBG_NONE && dbg < VKRT_DBG_RADIANCE)
  {
    switch(dbg)
    {
      case VKRT_DBG_METALLIC: b.radiance = f3(b.m.metallic); break;
      case VKRT_DBG_NORMAL: b.radiance = (b.fr.normal + f3(1.f)) * .5f; break;
      case VKRT_DBG_BASECOLOR: b.radiance = b.m.albedo; break;
      case VKRT_DBG_EMISSIVE: b.radiance = b.m.emission; break;
      case VKRT_DBG_ALPHA: b.radiance = f3(b.m.alpha); break;
      case VKRT_DBG_ROUGHNESS: b.radiance = f3(b.m.roughness); break;
      case VKRT_DBG_TEXCOORD: b.radiance = f3(texCoord.x, texCoord.y, 0.f); break;
      case VKRT_DBG_TANGENT: b.radiance = (b.fr.tangent + f3(1.f)) * .5f; break;
      default: b.radiance = f3(1000.f, 0.f, 0.f); break;
    }
    return false;
  }
  if(b.m.unlit)
  {
    b.radiance = b.radiance + b.m.albedo * b.throughput;
    return false;
  }
  if(dot(b.fr.normal, b.fr.ffnormal) > 0.0f)
    b.absorption = f3(0.0f);
  b.radiance += b.m.emission * b.throughput;
  b.throughput *= exp3(-b.absorption * b.hitT);
  return true;
}

// Phase B: DirectLight (pathtrace.glsl:97-188).
template <int PBR>
VKRT_D void bouncePhaseLight(const SceneView& sc, const FrameParams& fp, BounceState& b)
{
  const float3 V = -b.rayD;
  float3 lightContrib, lightDir;
  float  lightPdf;
  float  lightDist = 1e32f;
  bool   isLight   = false;
  b.visible  = false;
  b.vcontrib = f3(0.f);
  const float pSelectLight = fp.state.hdrMultiplier > 0.0f ? 0.5f : 1.0f;
  const int   nbLights     = sc.camera.nbLights;
  if(nbLights != 0 && rnd(b.seed) <= pSelectLight)
  {
    isLight = true;
    int lightIndex = (int)fminf(rnd(b.seed) * (float)nbLights, (float)nbLights);
    lightIndex     = min(lightIndex, nbLights - 1);  // memory safety only; unreachable for rand < 1
    const Light& light = sc.lights[lightIndex];
    const float3 ldir  = f3(light.direction[0], light.direction[1], light.direction[2]);
    float3 pointToLight = -ldir;
    float  rangeAtt = 1.0f, spotAtt = 1.0f;
    if(light.type != VKRT_LIGHT_DIRECTIONAL)
      pointToLight = f3(light.position[0], light.position[1], light.position[2]) - b.position;
    lightDist = length(pointToLight);
    if(light.type != VKRT_LIGHT_DIRECTIONAL)
      rangeAtt = rangeAttenuation(light.range, lightDist);
    if(light.type == VKRT_LIGHT_SPOT)
      spotAtt = spotAttenuation(pointToLight, ldir, light.outerConeCos, light.innerConeCos);
    lightContrib = rangeAtt * spotAtt * light.intensity * f3(light.color[0], light.color[1], light.color[2]);
    lightDir     = normalize(pointToLight);
    lightPdf     = 1.0f;
  }
  else
  {
    if(sc.sunsky.in_use == 1)
    {
      const float  sunRadius = (0.00465f * 10.0f) * sc.sunsky.sun_disk_scale;
      const float3 sd = f3(sc.sunsky.sun_direction[0], sc.sunsky.sun_direction[1], sc.sunsky.sun_direction[2]);
      float3 T, B;
      CreateCoordinateSystem(sd, T, B);
      float3 dir;
      dir.x    = rnd(b.seed) * sunRadius;
      dir.y    = rnd(b.seed) * sunRadius;
      dir.z    = sqrtf(fmaxf(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
      lightDir = normalize(T * dir.x + B * dir.y + sd * dir.z);
      lightContrib = sky::sunAndSky(sc.sunsky, lightDir);
      lightPdf = 0.5f;
    }
    else
    {
      float r0 = rnd(b.seed), r1 = rnd(b.seed), r2 = rnd(b.seed);
      lightContrib = environmentSample(sc, f3(r0, r1, r2), lightDir, lightPdf);
    }
    lightContrib *= fp.state.hdrMultiplier;
  }
  if(b.fr.isSubsurface || dot(lightDir, b.fr.ffnormal) > 0.0f)
  {
    float  bpdf = 0.f;
    float3 f = (PBR == 0) ? DisneyEval(b.m, b.fr, V, b.fr.ffnormal, lightDir, bpdf) : PbrEval(b.m, b.fr, V, b.fr.ffnormal, lightDir, bpdf);
    float  misWeight = isLight ? 1.0f : fmaxf(0.0f, powerHeuristic(lightPdf, bpdf));
    b.vcontrib = misWeight * f * fabsf(dot(lightDir, b.fr.ffnormal)) * lightContrib / lightPdf;
    b.visible  = true;
  }
  b.vcontrib *= b.throughput;
  b.lightDir  = lightDir;
  b.lightDist = lightDist;
}

// Phase C: BSDF sample, next ray, shadow ray, Russian roulette (pathtrace.glsl:281-340).
template <int PBR>
VKRT_D void bouncePhaseSample(const SceneView& sc, const FrameParams& fp, BounceState& b, ShadeOut& out)
{
  const int dbg      = fp.state.debugging_mode;
  const int depth    = fp.depth;
  const int maxDepth = fp.state.maxDepth;
  const float3 V = -b.rayD;
  float3 L   = f3(0.f);
  float  pdf = 0.f;
  float3 f   = (PBR == 0) ? DisneySample(b.m, b.fr, V, b.fr.ffnormal, L, pdf, b.seed) : PbrSample(b.m, b.fr, V, b.fr.ffnormal, L, pdf, b.seed);

  if(dot(b.fr.ffnormal, L) < 0.0f)
    b.absorption = -log3(b.m.attenuationColor) / f3(b.m.attenuationDistance);

  if(pdf > 0.0f)
    b.throughput *= f * fabsf(dot(b.fr.ffnormal, L)) / pdf;
  else
    return;  // break: no shadow ray, no RR draw

  if(dbg != VKRT_DBG_NONE && (depth == maxDepth - 1))
  {
    if(dbg == VKRT_DBG_RADIANCE) { b.radiance = b.vcontrib; return; }
    else if(dbg == VKRT_DBG_WEIGHT) { b.radiance = b.throughput; return; }
    else if(dbg == VKRT_DBG_RAYDIR) { b.radiance = (L + f3(1.f)) * 0.5f; return; }
  }

  const float rrPcont = fminf(maxcomp(b.throughput) * b.fr.eta * b.fr.eta + 0.001f, 0.95f);

  out.nextD = L;
  out.nextO = OffsetRay(b.position, dot(L, b.fr.ffnormal) > 0 ? b.fr.ffnormal : -b.fr.ffnormal);

  if(b.visible)
  {
    out.wantShadow = true;
    out.shO = out.nextO; out.shD = b.lightDir; out.shTmax = b.lightDist; out.shC = b.vcontrib; out.shSeed = b.seed;
    // RQ: HitTest draws on prd.seed during the shadow query; RTX: the shadow payload holds a copy
    if(sc.hasNonOpaque && fp.seedMode == VKRT_SEED_RAYQUERY)
      pcg(b.seed);
  }
  if(rnd(b.seed) >= rrPcont)
    return;
  b.throughput /= rrPcont;
  out.wantNext = (depth + 1 < maxDepth);
}

template <int PBR, int FM, bool HEAT>  // FM: distinct kernel symbols for the IEEE and the fast-math build of this file; HEAT: debug mode 12 (times every path)
__global__ void __launch_bounds__(kShadeBlock, kShadeMinBlocks) shadeKernel(const __grid_constant__ SceneView sc, const __grid_constant__ Wavefront wf,
                                                                            const __grid_constant__ FrameParams fp, TraceStats* stats)
{
  const uint32_t count = wf.counters[wf.cbase];
  const uint32_t lane  = threadIdx.x & 31u;
  for(uint32_t base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x)
  {
    const uint32_t i     = base + threadIdx.x;
    const bool     valid = i < count;
    const uint32_t t0    = HEAT ? deviceNs() : 0u;
    ShadeOut       out;
    out.wantNext = out.wantShadow = false;
    uint32_t    p = 0;
    BounceState b;
    bool        live = false;
    if(valid)
    {
      p = ldS(&wf.queueCur[i]);
      const float4 rd = ldS(&wf.rayD[p]), h = ldS(&wf.hit[p]);
      const float4 t4 = ldS(&wf.thr[p]), r4 = ldS(&wf.rad[p]), a4 = ldS(&wf.absb[p]);
      b.rayD = f3(rd); b.throughput = f3(t4); b.radiance = f3(r4); b.absorption = f3(a4);
      b.seed = __float_as_uint(t4.w);
      if(sc.hasNonOpaque)
        pcg(b.seed);  // the one draw of ClosestHit (see alphaDraw)
      live = bouncePhaseSurface(sc, fp, h, b);
    }
    __syncthreads();
    if(live)
      bouncePhaseLight<PBR>(sc, fp, b);
    __syncthreads();
    if(live)
      bouncePhaseSample<PBR>(sc, fp, b, out);
    if(valid)
    {
      stS(&wf.thr[p], f4(b.throughput, __uint_as_float(b.seed)));
      stS(&wf.rad[p], f4(b.radiance, 0.f));
      stS(&wf.absb[p], f4(b.absorption, 0.f));
      if(HEAT)
        wf.sum[p].w += (float)(deviceNs() - t0);  // heat map: the shading time of this bounce joins the traversal time of its rays
    }
    // warp-aggregated queue appends (all 32 lanes reach this point)
    const unsigned mNext = __ballot_sync(0xffffffffu, out.wantNext);
    const unsigned mSh   = __ballot_sync(0xffffffffu, out.wantShadow);
    uint32_t baseNext = 0, baseSh = 0;
    if(lane == 0)
    {
      if(mNext) baseNext = atomicAdd(&wf.counters[wf.cbase + 4], (uint32_t)__popc(mNext));  // the next bounce's queue
      if(mSh) baseSh = atomicAdd(&wf.counters[wf.cbase + 2], (uint32_t)__popc(mSh));
    }
    baseNext = __shfl_sync(0xffffffffu, baseNext, 0);
    baseSh   = __shfl_sync(0xffffffffu, baseSh, 0);
    const unsigned lt = (1u << lane) - 1u;
    if(out.wantNext)
    {
      stS(&wf.queueNext[baseNext + __popc(mNext & lt)], p);
      stS(&wf.rayO[p], f4(out.nextO, kInfinity));
      stS(&wf.rayD[p], f4(out.nextD, __uint_as_float(b.seed)));
    }
    if(out.wantShadow)
    {
      const uint32_t k = baseSh + __popc(mSh & lt);
      stS(&wf.shO[k], f4(out.shO, out.shTmax));
      stS(&wf.shD[k], f4(out.shD, __uint_as_float(out.shSeed)));
      stS(&wf.shC[k], f4(out.shC, __uint_as_float(p)));
    }
    __syncthreads();
  }
  if(stats && blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&stats->closestRays, (unsigned long long)count);
}

// ---- fused tail: the thin late bounces of a wavefront in ONE launch -------------------------------------------------------------
// From bounce `fp.depth` on, a path runs its own loop -- closest hit, the three shading phases, the shadow ray, the next bounce --
// like the reference's megakernel (pathtrace.glsl:193-343).  After three bounces a few percent of the paths are alive, and three
// launches per bounce cost more in launch gaps and kernel tails than in work (round 1: an interactive frame ran at half the
// batched rate).  Draw order, arithmetic and accumulation order are those of the wavefront kernels: images are bit-identical.
// Lanes are refilled one by one: a lane whose path has ended takes the next queued path at once (warp-aggregated fetch from the
// bounce's work cursor) instead of idling until the longest path of its warp is done -- under Russian roulette the remaining
// lengths of 32 paths differ widely.  VKRT_TAIL_REFILL=0 compiles the round's first version (one path per lane and loop trip).
#ifndef VKRT_TAIL_REFILL
#define VKRT_TAIL_REFILL 1
#endif
#ifndef VKRT_TAIL_MINBLOCKS
#define VKRT_TAIL_MINBLOCKS 4  // 128 registers, ~70 bytes of spills; 3 (162 registers, no spills) measured 1-2 % slower (profiles/README.md)
#endif
template <int PBR, int FM>
__global__ void __launch_bounds__(128, VKRT_TAIL_MINBLOCKS) tailKernel(const __grid_constant__ SceneView sc, const __grid_constant__ Wavefront wf,
                                                  const __grid_constant__ FrameParams fp0, TraceStats* stats)
{
  const uint32_t     count = wf.counters[wf.cbase];
  unsigned long long nClosest = 0, nShadow = 0;
#if VKRT_TAIL_REFILL
  uint32_t*      cursor = &wf.counters[wf.cbase + 1];  // this bounce's work cursor: zeroed by raygen, not used otherwise (no trace launch at this depth)
  const uint32_t lane   = threadIdx.x & 31u;
  const unsigned lt     = (1u << lane) - 1u;
  bool           active = false, drained = false;
  uint32_t       p = 0, raySeed = 0;
  float3         o = f3(0.f), d = f3(0.f);
  BounceState    b;
  FrameParams    fp = fp0;
  for(;;)
  {
    const unsigned idle = __ballot_sync(0xffffffffu, !active);
    if(idle && !drained)
    {
      uint32_t base = 0;
      if(lane == 0)
        base = atomicAdd(cursor, (uint32_t)__popc(idle));
      base    = __shfl_sync(0xffffffffu, base, 0);
      drained = base + (uint32_t)__popc(idle) >= count;
      const uint32_t i = base + (uint32_t)__popc(idle & lt);
      if(!active && i < count)
      {
        p = wf.queueCur[i];
        const float4 ro = wf.rayO[p], rd = wf.rayD[p], t4 = wf.thr[p], r4 = wf.rad[p], a4 = wf.absb[p];
        o = f3(ro); d = f3(rd);
        raySeed      = __float_as_uint(rd.w);
        b.throughput = f3(t4); b.radiance = f3(r4); b.absorption = f3(a4);
        b.seed       = __float_as_uint(t4.w);
        fp.depth     = fp0.depth;
        active       = true;
      }
    }
    if(!__any_sync(0xffffffffu, active))
      break;
    if(active)
    {
      // one bounce of this lane's path
      bool more = false;
      const TailHit h = traceSingle<false>(sc, o, d, kInfinity, raySeed);
      nClosest++;
      if(sc.hasNonOpaque)
        pcg(b.seed);  // the one draw of ClosestHit (see alphaDraw)
      b.rayD = d;
      const float4 hit = (h.tri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(h.t, h.u, h.v, __uint_as_float(h.tri));
      if(bouncePhaseSurface(sc, fp, hit, b))
      {
        bouncePhaseLight<PBR>(sc, fp, b);
        ShadeOut out;
        out.wantNext = out.wantShadow = false;
        bouncePhaseSample<PBR>(sc, fp, b, out);
        if(out.wantShadow)
        {
          nShadow++;
          const TailHit sh = traceSingle<true>(sc, out.shO, out.shD, out.shTmax, out.shSeed);
          if(sh.tri == ~0u)
            b.radiance = b.radiance + out.shC;  // the any-hit kernel's `rad += contribution`, same operand order
        }
        if(out.wantNext && fp.depth + 1 < fp0.state.maxDepth)
        {
          o = out.nextO; d = out.nextD; raySeed = b.seed;
          fp.depth++;
          more = true;
        }
      }
      if(!more)
      {
        wf.thr[p]  = f4(b.throughput, __uint_as_float(b.seed));
        wf.rad[p]  = f4(b.radiance, 0.f);
        wf.absb[p] = f4(b.absorption, 0.f);
        active     = false;
      }
    }
  }
#else
  for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
  {
    const uint32_t p  = wf.queueCur[i];
    const float4   ro = wf.rayO[p], rd = wf.rayD[p], t4 = wf.thr[p], r4 = wf.rad[p], a4 = wf.absb[p];
    float3         o = f3(ro), d = f3(rd);
    uint32_t       raySeed = __float_as_uint(rd.w);
    BounceState    b;
    b.throughput = f3(t4); b.radiance = f3(r4); b.absorption = f3(a4);
    b.seed       = __float_as_uint(t4.w);
    FrameParams fp = fp0;
    for(int depth = fp0.depth; depth < fp0.state.maxDepth; depth++)
    {
      fp.depth = depth;
      const TailHit h = traceSingle<false>(sc, o, d, kInfinity, raySeed);
      nClosest++;
      if(sc.hasNonOpaque)
        pcg(b.seed);  // the one draw of ClosestHit (see alphaDraw)
      b.rayD = d;
      const float4 hit = (h.tri == ~0u) ? f4(kInfinity, 0.f, 0.f, __uint_as_float(~0u)) : f4(h.t, h.u, h.v, __uint_as_float(h.tri));
      if(!bouncePhaseSurface(sc, fp, hit, b))
        break;
      bouncePhaseLight<PBR>(sc, fp, b);
      ShadeOut out;
      out.wantNext = out.wantShadow = false;
      bouncePhaseSample<PBR>(sc, fp, b, out);
      if(out.wantShadow)
      {
        nShadow++;
        const TailHit s = traceSingle<true>(sc, out.shO, out.shD, out.shTmax, out.shSeed);
        if(s.tri == ~0u)
          b.radiance = b.radiance + out.shC;  // the any-hit kernel's `rad += contribution`, same operand order
      }
      if(!out.wantNext)
        break;
      o = out.nextO; d = out.nextD; raySeed = b.seed;
    }
    wf.thr[p]  = f4(b.throughput, __uint_as_float(b.seed));
    wf.rad[p]  = f4(b.radiance, 0.f);
    wf.absb[p] = f4(b.absorption, 0.f);
  }
#endif
  if(stats)
  {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1)
    {
      nClosest += __shfl_xor_sync(0xffffffffu, nClosest, o);
      nShadow += __shfl_xor_sync(0xffffffffu, nShadow, o);
    }
    if((threadIdx.x & 31u) == 0 && (nClosest | nShadow))
    {
      atomicAdd(&stats->closestRays, nClosest);
      atomicAdd(&stats->shadowRays, nShadow);
    }
  }
}

#ifndef VKRT_FAST_MATH
// ---- sample finish: firefly clamp (pathtrace.glsl:380-384) + pixel sample sum (pathtrace.comp:102-105) ---
__global__ void __launch_bounds__(256) finishSampleKernel(const __grid_constant__ Wavefront wf, const __grid_constant__ FrameParams fp)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p >= wf.numPaths)
    return;
  float3 radiance = f3(wf.rad[p]);
  float  lum      = dot(radiance, f3(0.212671f, 0.715160f, 0.072169f));
  if(lum > fp.state.fireflyClampThreshold)
    radiance *= fp.state.fireflyClampThreshold / lum;
  float4 s  = wf.sum[p];
  wf.sum[p] = f4(s.x + radiance.x, s.y + radiance.y, s.z + radiance.z, s.w);  // .w: heat-map cost (trace kernels)
}

// ---- heat map: shaders/common.glsl:38-62 (fade, temperature) ----------------------------------------------------------
VKRT_D float fadeT(float low, float high, float value)
{
  const float mid = (low + high) * 0.5f, range = (high - low) * 0.5f;
  const float x   = 1.0f - clampf(fabsf(mid - value) / range, 0.0f, 1.0f);
  return smoothstepf(0.0f, 1.0f, x);
}
VKRT_D float3 temperature(float intensity)
{
  return f3(0.f, 0.f, 1.f) * fadeT(-0.25f, 0.25f, intensity) + f3(0.f, 1.f, 1.f) * fadeT(0.0f, 0.5f, intensity)
         + f3(0.f, 1.f, 0.f) * fadeT(0.25f, 0.75f, intensity) + f3(1.f, 1.f, 0.f) * fadeT(0.5f, 1.0f, intensity)
         + f3(1.f, 0.f, 0.f) * smoothstepf(0.75f, 1.0f, intensity);
}
// The reference colours a pixel by the nanoseconds its invocation took (clockRealtimeEXT around the sample loop,
// pathtrace.comp:89,108-117).  The wavefront equivalent: every kernel that works on a path adds the nanoseconds (%globaltimer) its
// lane held that path -- traversal of its closest and shadow rays (waiting for the rest of the warp included, as a megakernel
// invocation waits for its neighbours) and the shading of each bounce -- to the path's cost, and the accumulate kernel maps the sum
// through the reference's temperature() and min/maxHeatmap.  Like any clock it is not reproducible from run to run.

// ---- accumulate: pathtrace.comp:106,122-133 (running mean, frame order) ------------------------------------
__global__ void __launch_bounds__(256) accumulateKernel(const __grid_constant__ Wavefront wf, const __grid_constant__ FrameParams fp,
                                                        float4* __restrict__ image, float4* __restrict__ mirror, bool fromRadiance)
{
  const uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
  if(lp >= wf.numLocalPixels)
    return;
  const uint32_t pixel = wf.pixelOfLocal[lp];
  float4         px    = image[pixel];
  float3         acc   = f3(px);
  for(int f = 0; f < fp.framesInBatch; f++)
  {
    const int frame      = fp.state.frame + f;
    float4 sum;
    if(fromRadiance)
    {
      // one sample per frame: finishSampleKernel's work, same operations (firefly clamp, 0 + radiance as the sample sum)
      float3 radiance = f3(wf.rad[(size_t)f * wf.numLocalPixels + lp]);
      float  lum      = dot(radiance, f3(0.212671f, 0.715160f, 0.072169f));
      if(lum > fp.state.fireflyClampThreshold)
        radiance *= fp.state.fireflyClampThreshold / lum;
      sum = f4(0.0f + radiance.x, 0.0f + radiance.y, 0.0f + radiance.z, 0.0f);
    }
    else
      sum = wf.sum[(size_t)f * wf.numLocalPixels + lp];
    float3       pixelColor = f3(sum);
    pixelColor /= (float)fp.state.maxSamples;
    if(fp.state.debugging_mode == VKRT_DBG_HEATMAP)
    {
      const float low = (float)fp.state.minHeatmap, high = (float)fp.state.maxHeatmap;
      pixelColor      = temperature(clampf((sum.w - low) / (high - low), 0.0f, 1.0f));
    }
    if(frame > 0)
      acc = mix(acc, pixelColor, 1.0f / (float)(frame + 1));
    else
      acc = pixelColor;
  }
  image[pixel] = f4(acc, 1.f);
  if(mirror)
    mirror[pixel] = f4(acc, 1.f);  // tile shards: the root's image, mapped from another process (vkrt_output_set_mirror)
}

// ---- tonemap: post.frag:98-147 (autoExposure off), tonemapping.glsl:24-65 -------------------------------------
VKRT_D float3 linearTosRGB(float3 c) { return pow3(c, 1.0f / 2.2f); }
VKRT_D float3 sRGBToLinear(float3 c) { return pow3(c, 2.2f); }
VKRT_D float3 uncharted2Impl(float3 color)
{
  const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
  return ((color * (A * color + C * B) + D * E) / (color * (A * color + B) + D * F)) - E / F;
}
VKRT_D float3 toneMapUncharted(float3 color)
{
  const float W   = 11.2f;
  color           = uncharted2Impl(color * 2.0f);
  float3 whiteScale = f3(1.0f) / uncharted2Impl(f3(W));
  return linearTosRGB(color * whiteScale);
}
// One level of RenderOutput::genMipmap (src/render_output.cpp:188-193): vkCmdBlitImage with VK_FILTER_LINEAR from level k to
// level k+1 (extent max(1, size >> 1)): source sampled at u = (x + 0.5) * srcW / dstW, linear, clamp-to-edge.
__global__ void __launch_bounds__(256) mipDownKernel(const float4* __restrict__ src, int sw, int sh, float4* __restrict__ dst, int dw, int dh)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= dw * dh)
    return;
  const int   x = i % dw, y = i / dw;
  const float u = ((float)x + 0.5f) * ((float)sw / (float)dw) - 0.5f, v = ((float)y + 0.5f) * ((float)sh / (float)dh) - 0.5f;
  const float x0f = floorf(u), y0f = floorf(v), ax = u - x0f, ay = v - y0f;
  const int   x0 = min(sw - 1, max(0, (int)x0f)), x1 = min(sw - 1, max(0, (int)x0f + 1));
  const int   y0 = min(sh - 1, max(0, (int)y0f)), y1 = min(sh - 1, max(0, (int)y0f + 1));
  const float4 c00 = src[(size_t)y0 * sw + x0], c10 = src[(size_t)y0 * sw + x1], c01 = src[(size_t)y1 * sw + x0], c11 = src[(size_t)y1 * sw + x1];
  const float4 top = c00 * (1.0f - ax) + c10 * ax, bot = c01 * (1.0f - ax) + c11 * ax;
  dst[i] = top * (1.0f - ay) + bot * ay;
}

// NEAREST fetch of a mip level (default VkSamplerCreateInfo of the post pass, src/render_output.cpp:96-98)
VKRT_D float3 mipFetch(const MipView& m, float2 uv, int lod)
{
  lod = min(max(lod, 0), m.count - 1);
  const int w = m.w[lod], h = m.h[lod];
  const int x = min(w - 1, max(0, (int)floorf(uv.x * (float)w))), y = min(h - 1, max(0, (int)floorf(uv.y * (float)h)));
  return f3(m.level[lod][(size_t)y * w + x]);
}
VKRT_D float postLuminance(float3 c) { return dot(c, f3(0.2126f, 0.7152f, 0.0722f)); }
VKRT_D float xyzY(float3 rgb) { return (0.3575761f * rgb.x + 0.7151522f * rgb.y) + 0.1191920f * rgb.z; }  // post.frag:57 (column-major quirk)

__global__ void __launch_bounds__(256) tonemapKernel(const float4* __restrict__ image, int W, int H, const __grid_constant__ Tonemapper tm,
                                                     const __grid_constant__ MipView mips, float4* __restrict__ outF, uchar4* __restrict__ out8)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= W * H)
    return;
  const int x = i % W, y = i / W;
  const float2 uvCoords = f2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
  const int sx = min(W - 1, max(0, (int)floorf(uvCoords.x * tm.zoom * (float)W)));
  const int sy = min(H - 1, max(0, (int)floorf(uvCoords.y * tm.zoom * (float)H)));
  const float4 hdr = image[(size_t)sy * W + sx];
  float3 hdrRgb = f3(hdr);
  if((tm.autoExposure & 1) && mips.count > 0)
  {
    // post.frag:101-111: average = the 1x1 top mip; toneExposure (post.frag:64-70) or toneLocalExposure (:72-95)
    const float logAvgLum = postLuminance(mipFetch(mips, f2(0.5f, 0.5f), 20));
    const float XYZy      = xyzY(hdrRgb);
    const float Y         = (tm.key / logAvgLum) * XYZy;
    float       Yd;
    if((tm.autoExposure >> 1) & 1)
    {
      const float2 uvz    = f2(uvCoords.x * tm.zoom, uvCoords.y * tm.zoom);
      const float  factor = tm.key / logAvgLum;
      float        La = 0.f, sc = 1.f;
      for(int l = 0; l < 7; ++l)
      {
        const float v1 = postLuminance(mipFetch(mips, uvz, l)) * factor;
        const float v2 = postLuminance(mipFetch(mips, uvz, l + 1)) * factor;
        if(fabsf(v1 - v2) / ((tm.key * 4.0f / (sc * sc)) + v1) > 0.05f)  // 4 = pow(2, phi), phi = 2
        {
          La = v1;
          break;
        }
        else
          La = v2;
        sc *= 2.f;
      }
      Yd = Y / (1.0f + La);
    }
    else
      Yd = (Y * (1.0f + Y / (tm.Ywhite * tm.Ywhite))) / (1.0f + Y);
    hdrRgb = hdrRgb / XYZy * Yd;
  }
  float3 color = toneMapUncharted(hdrRgb * tm.avgLum);
  if(tm.dither > 0)
  {
    uint32_t rx = (uint32_t)x, ry = (uint32_t)y, rz = 0u;
    pcg3d(rx, ry, rz);
    const float3 noise = f3(u2f(0x3f800000u | (rx >> 9)) - 1.0f, u2f(0x3f800000u | (ry >> 9)) - 1.0f, u2f(0x3f800000u | (rz >> 9)) - 1.0f);
    const float  quant = 1.f / 255.f;
    const float3 lin   = sRGBToLinear(color);
    const float3 c0    = floor3(linearTosRGB(lin) / quant) * quant;
    const float3 c1    = c0 + quant;
    const float3 discr = mix(sRGBToLinear(c0), sRGBToLinear(c1), noise);
    color = f3(discr.x < lin.x ? c1.x : c0.x, discr.y < lin.y ? c1.y : c0.y, discr.z < lin.z ? c1.z : c0.z);
  }
  color = clamp3(mix(f3(0.5f), color, tm.contrast), 0.f, 1.f);
  color = pow3(color, 1.0f / tm.brightness);
  const float3 gray = f3(dot(color, f3(0.299f, 0.587f, 0.114f)));
  color = mix(gray, color, tm.saturation);
  const float2 uv = f2(((uvCoords.x * tm.renderingRatio[0]) - 0.5f) * 2.0f, ((uvCoords.y * tm.renderingRatio[1]) - 0.5f) * 2.0f);
  color *= 1.0f - dot(uv, uv) * tm.vignette;
  if(outF)
    outF[i] = f4(color, hdr.w);
  if(out8)
  {
    // UNORM8 colour attachment write: clamp, scale, round to nearest even
    auto q = [](float f) { return (unsigned char)__float2int_rn(fminf(fmaxf(f, 0.f), 1.f) * 255.0f); };
    out8[i] = make_uchar4(q(color.x), q(color.y), q(color.z), q(hdr.w));
  }
}

#endif  // !VKRT_FAST_MATH

// ---- launchers ---------------------------------------------------------------------------------------------
#ifdef VKRT_FAST_MATH
#define VKRT_SHADE_FM 1
#define VKRT_LAUNCH_SHADE launchShadeFast
#define VKRT_LAUNCH_TAIL launchTailFast
#else
#define VKRT_SHADE_FM 0
#define VKRT_LAUNCH_SHADE launchShadePrecise
#define VKRT_LAUNCH_TAIL launchTailPrecise
void launchRaygen(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, int)
{
  raygenKernel<<<(wf.numPaths + 255) / 256, 256, 0, s>>>(scene, wf, fp);
}
#endif
template <int PBR, bool HEAT>
static void launchShadeT(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs)
{
  static int perSM = 0;
  if(!perSM)
  {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, shadeKernel<PBR, VKRT_SHADE_FM, HEAT>, kShadeBlock, 0);
    if(perSM < 1) perSM = 1;
  }
  int grid = numSMs * perSM;
  int need = (int)((wf.numPaths + kShadeBlock - 1) / kShadeBlock);
  if(grid > need) grid = need;
  if(grid < 1) grid = 1;
  shadeKernel<PBR, VKRT_SHADE_FM, HEAT><<<grid, kShadeBlock, 0, s>>>(scene, wf, fp, stats);
}
void VKRT_LAUNCH_SHADE(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs)
{
  // one kernel per BSDF model (RtxState.pbrMode, pathtrace.glsl:40-56): halves the code and register pressure
  const bool heat = fp.state.debugging_mode == VKRT_DBG_HEATMAP;  // the timed variant (debug mode 12) is a kernel of its own
  if(fp.state.pbrMode == 0)
    heat ? launchShadeT<0, true>(s, scene, wf, fp, stats, numSMs) : launchShadeT<0, false>(s, scene, wf, fp, stats, numSMs);
  else
    heat ? launchShadeT<1, true>(s, scene, wf, fp, stats, numSMs) : launchShadeT<1, false>(s, scene, wf, fp, stats, numSMs);
}
template <int PBR>
static void launchTailT(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs)
{
  static int perSM = 0;
  if(!perSM)
  {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, tailKernel<PBR, VKRT_SHADE_FM>, 128, 0);
    if(perSM < 1) perSM = 1;
  }
  tailKernel<PBR, VKRT_SHADE_FM><<<numSMs * perSM, 128, 0, s>>>(scene, wf, fp, stats);
}
void VKRT_LAUNCH_TAIL(cudaStream_t s, const SceneView& scene, const Wavefront& wf, const FrameParams& fp, TraceStats* stats, int numSMs)
{
  if(fp.state.pbrMode == 0)
    launchTailT<0>(s, scene, wf, fp, stats, numSMs);
  else
    launchTailT<1>(s, scene, wf, fp, stats, numSMs);
}
#ifndef VKRT_FAST_MATH
// tile shards: image[pixels[i]] <-> compact[i] (multi-GPU exchange of owned pixels instead of a full-image reduce)
__global__ void __launch_bounds__(256) shardCopyKernel(float4* __restrict__ image, const uint32_t* __restrict__ pixels, uint32_t n,
                                                       float4* __restrict__ compact, bool toImage)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  if(toImage)
    image[pixels[i]] = compact[i];
  else
    compact[i] = image[pixels[i]];
}
void launchShardCopy(cudaStream_t s, float4* image, const uint32_t* pixels, uint32_t n, float4* compact, bool toImage)
{
  if(n)
    shardCopyKernel<<<(n + 255) / 256, 256, 0, s>>>(image, pixels, n, compact, toImage);
}
void launchFinishSample(cudaStream_t s, const Wavefront& wf, const FrameParams& fp, int)
{
  finishSampleKernel<<<(wf.numPaths + 255) / 256, 256, 0, s>>>(wf, fp);
}
void launchAccumulate(cudaStream_t s, const Wavefront& wf, const FrameParams& fp, float4* image, float4* mirror, int, bool fromRadiance)
{
  accumulateKernel<<<(wf.numLocalPixels + 255) / 256, 256, 0, s>>>(wf, fp, image, mirror, fromRadiance);
}
void launchMipDown(cudaStream_t s, const float4* src, int sw, int sh, float4* dst, int dw, int dh)
{
  mipDownKernel<<<(dw * dh + 255) / 256, 256, 0, s>>>(src, sw, sh, dst, dw, dh);
}
void launchTonemap(cudaStream_t s, const float4* image, int w, int h, const Tonemapper& tm, const MipView& mips, float4* outF, uchar4* out8)
{
  tonemapKernel<<<(w * h + 255) / 256, 256, 0, s>>>(image, w, h, tm, mips, outF, out8);
}

#endif  // !VKRT_FAST_MATH

}  // namespace vkrt
