// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or imported by the product.
//
// o_pathtrace.h: CPU restatement (megakernel form, one call per pixel sample) of
//   shade state        shaders/shade_state.glsl:63-145
//   material resolve   shaders/gltf_material.glsl:37-193
//   env sampling       shaders/env_sampling.glsl:38-135
//   punctual lights    shaders/punctual.glsl:28-51
//   integrator         shaders/pathtrace.glsl:40-387
//   pixel loop         shaders/pathtrace.comp:87-134 (RQ seed) / pathtrace.rgen:60-108 (RTX seed)
#pragma once
#include "o_scene.h"
#include "o_bsdf.h"
#include "o_sunsky.h"

namespace orc {

struct Ray
{
  vec3 origin, direction;
};

// shaders/globals.glsl:53-63 -- the matrices are looked up from the node instead of copied
struct PtPayload
{
  uint32_t seed;
  float    hitT;
  int      primitiveID, instanceID, instanceCustomIndex;
  vec2     baryCoord;
};

struct ShadeState
{
  vec3     normal, geom_normal, position;
  vec2     text_coords;
  vec3     tangent_u, tangent_v, color;
  uint32_t matIndex;
};

struct RenderCounters
{
  uint64_t      closestRays = 0, shadowRays = 0;
  TraceCounters trace;
};

struct PathCtx
{
  const Scene*    scene;
  RtxState        rtx;
  int             seedMode;  // 0 RQ, 1 RTX
  PtPayload       prd;
  RenderCounters* cnt;
};

// ---- shaders/shade_state.glsl:63-145 ---------------------------------------------
inline ShadeState GetShadeState(const Scene& s, const PtPayload& h)
{
  ShadeState      ss;
  const Node&     nd = s.nodes[h.instanceID];
  const PrimMesh& pm = s.meshes[h.instanceCustomIndex];
  const vec3      bary = V3(1.0f - h.baryCoord.x - h.baryCoord.y, h.baryCoord.x, h.baryCoord.y);

  const VertexAttributes& a0 = pm.vtx[pm.idx[h.primitiveID * 3 + 0]];
  const VertexAttributes& a1 = pm.vtx[pm.idx[h.primitiveID * 3 + 1]];
  const VertexAttributes& a2 = pm.vtx[pm.idx[h.primitiveID * 3 + 2]];
  const uint32_t          matIndex = (uint32_t)std::max(0, pm.material);

  const vec3 pos0 = V3(a0.position[0], a0.position[1], a0.position[2]);
  const vec3 pos1 = V3(a1.position[0], a1.position[1], a1.position[2]);
  const vec3 pos2 = V3(a2.position[0], a2.position[1], a2.position[2]);
  const vec3 position       = pos0 * bary.x + pos1 * bary.y + pos2 * bary.z;
  const vec3 world_position = xformPoint(nd.world, position);

  vec3 nrm0         = decompress_unit_vec(a0.normal);
  vec3 nrm1         = decompress_unit_vec(a1.normal);
  vec3 nrm2         = decompress_unit_vec(a2.normal);
  vec3 normal       = normalize(nrm0 * bary.x + nrm1 * bary.y + nrm2 * bary.z);
  vec3 world_normal = normalize(xformNormalT(nd.worldInv, normal));
  vec3 geom_normal  = normalize(cross(pos1 - pos0, pos2 - pos0));
  vec3 wgeom_normal = normalize(xformNormalT(nd.worldInv, geom_normal));

  float h0 = (floatBitsToInt(a0.texcoord[1]) & 1) == 1 ? 1.0f : -1.0f;
  vec3  tng0 = decompress_unit_vec(a0.tangent);
  vec3  tng1 = decompress_unit_vec(a1.tangent);
  vec3  tng2 = decompress_unit_vec(a2.tangent);
  vec3  tangent = (tng0 * bary.x + tng1 * bary.y + tng2 * bary.z);
  tangent            = normalize(tangent);
  vec3 world_tangent = normalize(xformVector(nd.world, tangent));
  world_tangent      = normalize(world_tangent - dot(world_tangent, world_normal) * world_normal);
  vec3 world_binormal = cross(world_normal, world_tangent) * h0;

  auto decode_texture = [](const float* t) { return V2(t[0], uintBitsToFloat(floatBitsToUint(t[1]) & ~1u)); };
  const vec2 uv0 = decode_texture(a0.texcoord), uv1 = decode_texture(a1.texcoord), uv2 = decode_texture(a2.texcoord);
  const vec2 texcoord0 = uv0 * bary.x + uv1 * bary.y + uv2 * bary.z;

  const vec4 col0 = unpackUnorm4x8(a0.color), col1 = unpackUnorm4x8(a1.color), col2 = unpackUnorm4x8(a2.color);
  const vec4 color = col0 * bary.x + col1 * bary.y + col2 * bary.z;

  ss.normal      = world_normal;
  ss.geom_normal = wgeom_normal;
  ss.position    = world_position;
  ss.text_coords = texcoord0;
  ss.tangent_u   = world_tangent;
  ss.tangent_v   = world_binormal;
  ss.color       = xyz(color);
  ss.matIndex    = matIndex;
  if(dot(ss.normal, ss.geom_normal) <= 0)
    ss.normal *= -1.0f;
  return ss;
}

// ---- shaders/gltf_material.glsl ------------------------------------------------------
// gltf_material.glsl:37-46 (SRGB_FAST_APPROXIMATION)
inline vec4 SRGBtoLINEAR(vec4 c) { return V4(pow3(xyz(c), V3(2.2f)), c.w); }

inline vec4 tex(const Scene& s, int idx, vec2 uv) { return sampleTexture(s.textures[idx], uv); }

// gltf_material.glsl:52-93
inline void GetMetallicRoughness(const Scene& s, State& state, const GltfShadeMaterial& material)
{
  float dielectricSpecular = (material.ior - 1) / (material.ior + 1);
  dielectricSpecular *= dielectricSpecular;
  float perceptualRoughness = material.pbrRoughnessFactor;
  float metallic            = material.pbrMetallicFactor;
  if(material.pbrMetallicRoughnessTexture > -1)
  {
    vec4 mrSample       = tex(s, material.pbrMetallicRoughnessTexture, state.texCoord);
    perceptualRoughness = mrSample.y * perceptualRoughness;
    metallic            = mrSample.z * metallic;
  }
  vec4 baseColor = V4(material.pbrBaseColorFactor[0], material.pbrBaseColorFactor[1], material.pbrBaseColorFactor[2], material.pbrBaseColorFactor[3]);
  if(material.pbrBaseColorTexture > -1)
    baseColor = baseColor * SRGBtoLINEAR(tex(s, material.pbrBaseColorTexture, state.texCoord));
  vec3 f0             = mix(V3(dielectricSpecular), xyz(baseColor), metallic);
  state.mat.albedo    = xyz(baseColor);
  state.mat.metallic  = metallic;
  state.mat.roughness = perceptualRoughness;
  state.mat.f0        = f0;
  state.mat.alpha     = baseColor.w;
}

// gltf_material.glsl:104-193
inline void GetMaterialsAndTextures(const Scene& s, State& state, const Ray& r)
{
  const GltfShadeMaterial& material = s.materials[state.matID];
  state.mat.specular     = 0.5f;
  state.mat.subsurface   = 0;
  state.mat.specularTint = 1;
  state.mat.sheen        = 0;
  state.mat.sheenTint    = V3(0);

  const float* M = material.uvTransform;
  vec2         tc = state.texCoord;
  state.texCoord  = V2(((tc.x * M[0] + tc.y * M[1]) + M[2]) + M[3], ((tc.x * M[4] + tc.y * M[5]) + M[6]) + M[7]);
  // mat3 TBN = mat3(tangent, bitangent, normal) captured before the normal map
  const vec3 T0 = state.tangent, B0 = state.bitangent, N0 = state.normal;
  auto TBNmul = [&](vec3 v) { return T0 * v.x + B0 * v.y + N0 * v.z; };

  if(material.normalTexture > -1)
  {
    vec3 normalVector = xyz(tex(s, material.normalTexture, state.texCoord));
    normalVector      = normalize(normalVector * 2.0f - 1.0f);
    normalVector      = normalVector * V3(material.normalTextureScale, material.normalTextureScale, 1.0f);
    state.normal      = normalize(TBNmul(normalVector));
    state.ffnormal    = dot(state.normal, r.direction) <= 0.0f ? state.normal : -state.normal;
    CreateCoordinateSystem(state.ffnormal, state.tangent, state.bitangent);
  }

  state.mat.emission = V3(material.emissiveFactor[0], material.emissiveFactor[1], material.emissiveFactor[2]);
  if(material.emissiveTexture > -1)
    state.mat.emission *= xyz(SRGBtoLINEAR(tex(s, material.emissiveTexture, state.texCoord)));

  GetMetallicRoughness(s, state, material);
  state.mat.roughness = std::fmax(state.mat.roughness, 0.001f);

  state.mat.transmission = material.transmissionFactor;
  if(material.transmissionTexture > -1)
    state.mat.transmission *= tex(s, material.transmissionTexture, state.texCoord).x;

  state.mat.ior = material.ior;
  state.eta     = dot(state.normal, state.ffnormal) > 0.0f ? (1.0f / state.mat.ior) : state.mat.ior;

  state.mat.unlit = (material.unlit == 1);

  state.mat.anisotropy = material.anisotropy;
  float aspect         = std::sqrt(1.0f - material.anisotropy * 0.9f);
  state.mat.ax         = std::fmax(0.001f, state.mat.roughness / aspect);
  state.mat.ay         = std::fmax(0.001f, state.mat.roughness * aspect);

  if(material.anisotropy > 0)
  {
    state.tangent   = normalize(TBNmul(V3(material.anisotropyDirection[0], material.anisotropyDirection[1], material.anisotropyDirection[2])));
    state.bitangent = normalize(cross(state.normal, state.tangent));
  }

  state.mat.attenuationColor    = V3(material.attenuationColor[0], material.attenuationColor[1], material.attenuationColor[2]);
  state.mat.attenuationDistance = material.attenuationDistance;
  state.mat.thinwalled          = material.thicknessFactor == 0;

  state.mat.clearcoat          = material.clearcoatFactor;
  state.mat.clearcoatRoughness = material.clearcoatRoughness;
  if(material.clearcoatTexture > -1)
    state.mat.clearcoat *= tex(s, material.clearcoatTexture, state.texCoord).x;
  if(material.clearcoatRoughnessTexture > -1)
    state.mat.clearcoatRoughness *= tex(s, material.clearcoatRoughnessTexture, state.texCoord).y;
  state.mat.clearcoatRoughness = std::fmax(state.mat.clearcoatRoughness, 0.001f);

  vec4 sheen          = unpackUnorm4x8(material.sheen);
  state.mat.sheenTint = xyz(sheen);
  state.mat.sheen     = sheen.w;
}

// ---- shaders/punctual.glsl:28-51 --------------------------------------------------
inline float getRangeAttenuation(float range, float distance)
{
  if(range <= 0.0f)
    return 1.0f;
  return std::fmax(std::fmin(1.0f - std::pow(distance / range, 4.0f), 1.0f), 0.0f) / std::pow(distance, 2.0f);
}
inline float getSpotAttenuation(vec3 pointToLight, vec3 spotDirection, float outerConeCos, float innerConeCos)
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

// ---- shaders/env_sampling.glsl:38-99 ----------------------------------------------
inline vec3 Environment_sample(const Scene& s, vec3 xi, vec3& to_light, float& pdf)
{
  uint32_t       width = s.envW, height = s.envH;
  const uint32_t size = width * height;
  const uint32_t idx  = std::min(uint32_t(xi.x * float(size)), size - 1);
  EnvAccel       sd   = s.envAccel[idx];
  uint32_t       env_idx;
  if(xi.y < sd.q)
  {
    env_idx = idx;
    xi.y /= sd.q;
    pdf = sd.pdf;
  }
  else
  {
    env_idx = sd.alias;
    xi.y    = (xi.y - sd.q) / (1.0f - sd.q);
    pdf     = sd.aliasPdf;
  }
  const uint32_t px = env_idx % width;
  uint32_t       py = env_idx / width;
  const float    u       = float(px + xi.y) / float(width);
  const float    phi     = u * (2.0f * K_PI) - K_PI;
  float          sin_phi = std::sin(phi);
  float          cos_phi = std::cos(phi);
  const float    step_theta = K_PI / float(height);
  const float    theta0     = float(py) * step_theta;
  const float    cos_theta  = std::cos(theta0) * (1.0f - xi.z) + std::cos(theta0 + step_theta) * xi.z;
  const float    theta      = std::acos(cos_theta);
  const float    sin_theta  = std::sin(theta);
  const float    v          = theta * K_1_OVER_PI;
  to_light = V3(cos_phi * sin_theta, cos_theta, sin_phi * sin_theta);
  return sampleEnv(s, V2(u, v));
}

// ---- shaders/env_sampling.glsl:105-135 --------------------------------------------
inline vec4 EnvSample(PathCtx& c, vec3& radiance)
{
  const Scene& s = *c.scene;
  vec3         lightDir;
  float        pdf;
  if(s.sunsky.in_use == 1)
  {
    float sun_radius = (0.00465f * 10.0f) * s.sunsky.sun_disk_scale;
    vec3  sd = V3(s.sunsky.sun_direction[0], s.sunsky.sun_direction[1], s.sunsky.sun_direction[2]);
    vec3  T, B;
    CreateCoordinateSystem(sd, T, B);
    vec3 dir;
    dir.x    = rnd(c.prd.seed) * sun_radius;
    dir.y    = rnd(c.prd.seed) * sun_radius;
    dir.z    = std::sqrt(std::fmax(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
    lightDir = normalize(T * dir.x + B * dir.y + sd * dir.z);
    radiance = sky::sun_and_sky(s.sunsky, lightDir);
    pdf      = 0.5f;
  }
  else
  {
    float a = rnd(c.prd.seed), b = rnd(c.prd.seed), d = rnd(c.prd.seed);
    radiance = Environment_sample(s, V3(a, b, d), lightDir, pdf);
  }
  radiance *= c.rtx.hdrMultiplier;
  return V4(lightDir, pdf);
}

// ---- trace wrappers: shaders/traceray_rq.glsl:108-185 -------------------------------
inline void ClosestHit(PathCtx& c, const Ray& r)
{
  const Scene& s = *c.scene;
  c.prd.hitT     = INFINITY_RT;
  Hit h          = traceRay(s, r.origin, r.direction, INFINITY_RT, c.prd.seed, true, c.cnt ? &c.cnt->trace : nullptr);
  if(c.cnt) c.cnt->closestRays++;
  if(s.hasNonOpaque)
    pcg(c.prd.seed);  // one draw per trace call (see alphaDraw)
  if(h.tri != ~0u)
  {
    const Tri& T              = s.tris[h.tri];
    c.prd.hitT                = h.t;
    c.prd.primitiveID         = (int)T.prim;
    c.prd.instanceID          = (int)T.node;
    c.prd.instanceCustomIndex = (int)s.nodes[T.node].primMesh;
    c.prd.baryCoord           = V2(h.u, h.v);
  }
}
inline bool AnyHit(PathCtx& c, const Ray& r, float maxDist)
{
  const Scene& s = *c.scene;
  Hit h = traceRay(s, r.origin, r.direction, maxDist, c.prd.seed, false, c.cnt ? &c.cnt->trace : nullptr);
  if(c.cnt) c.cnt->shadowRays++;
  // RQ: HitTest draws on prd.seed during shadow rays too; RTX: the shadow payload carries a copy
  // (shaders/traceray_rtx.glsl:54-55), so the path seed does not move.
  if(s.hasNonOpaque && c.seedMode == 0)
    pcg(c.prd.seed);
  return h.tri != ~0u;
}

// ---- shaders/pathtrace.glsl:40-56 --------------------------------------------------
inline vec3 Eval(PathCtx& c, const State& state, vec3 V, vec3 N, vec3 L, float& pdf)
{
  return c.rtx.pbrMode == 0 ? DisneyEval(state, V, N, L, pdf) : PbrEval(state, V, N, L, pdf);
}
inline vec3 Sample(PathCtx& c, State& state, vec3 V, vec3 N, vec3& L, float& pdf, uint32_t& seed)
{
  return c.rtx.pbrMode == 0 ? DisneySample(state, V, N, L, pdf, seed) : PbrSample(state, V, N, L, pdf, seed);
}
// pathtrace.glsl:61-83
inline vec3 DebugInfo(PathCtx& c, const State& state)
{
  switch(c.rtx.debugging_mode)
  {
    case VKRT_DBG_METALLIC: return V3(state.mat.metallic);
    case VKRT_DBG_NORMAL: return (state.normal + V3(1)) * .5f;
    case VKRT_DBG_BASECOLOR: return state.mat.albedo;
    case VKRT_DBG_EMISSIVE: return state.mat.emission;
    case VKRT_DBG_ALPHA: return V3(state.mat.alpha);
    case VKRT_DBG_ROUGHNESS: return V3(state.mat.roughness);
    case VKRT_DBG_TEXCOORD: return V3(state.texCoord.x, state.texCoord.y, 0);
    case VKRT_DBG_TANGENT: return (state.tangent + V3(1)) * .5f;
  }
  return V3(1000, 0, 0);
}

struct VisibilityContribution
{
  vec3  radiance, lightDir;
  float lightDist;
  bool  visible;
};

// pathtrace.glsl:97-188
inline VisibilityContribution DirectLight(PathCtx& c, const Ray& r, const State& state)
{
  const Scene& s = *c.scene;
  vec3  Li = V3(0);
  float lightPdf;
  vec3  lightContrib;
  vec3  lightDir;
  float lightDist = 1e32f;
  bool  isLight   = false;

  VisibilityContribution contrib;
  contrib.radiance  = V3(0);
  contrib.visible   = false;
  contrib.lightDir  = V3(0);
  contrib.lightDist = 0;

  float p_select_light = c.rtx.hdrMultiplier > 0.0f ? 0.5f : 1.0f;

  if(s.camera.nbLights != 0 && rnd(c.prd.seed) <= p_select_light)
  {
    isLight = true;
    int          light_index = int(std::fmin(rnd(c.prd.seed) * s.camera.nbLights, (float)s.camera.nbLights));
    const Light& light = s.lights[light_index];
    vec3  ldir = V3(light.direction[0], light.direction[1], light.direction[2]);
    vec3  pointToLight     = -ldir;
    float rangeAttenuation = 1.0f;
    float spotAttenuation  = 1.0f;
    if(light.type != VKRT_LIGHT_DIRECTIONAL)
      pointToLight = V3(light.position[0], light.position[1], light.position[2]) - state.position;
    lightDist = length(pointToLight);
    if(light.type != VKRT_LIGHT_DIRECTIONAL)
      rangeAttenuation = getRangeAttenuation(light.range, lightDist);
    if(light.type == VKRT_LIGHT_SPOT)
      spotAttenuation = getSpotAttenuation(pointToLight, ldir, light.outerConeCos, light.innerConeCos);
    vec3 intensity = rangeAttenuation * spotAttenuation * light.intensity * V3(light.color[0], light.color[1], light.color[2]);
    lightContrib = intensity;
    lightDir     = normalize(pointToLight);
    lightPdf     = 1.0f;
  }
  else
  {
    vec4 dirPdf = EnvSample(c, lightContrib);
    lightDir    = xyz(dirPdf);
    lightPdf    = dirPdf.w;
  }

  if(state.isSubsurface || dot(lightDir, state.ffnormal) > 0.0f)
  {
    float bpdf      = 0.f;
    vec3  f         = Eval(c, state, -r.direction, state.ffnormal, lightDir, bpdf);
    float misWeight = isLight ? 1.0f : std::fmax(0.0f, powerHeuristic(lightPdf, bpdf));
    Li += misWeight * f * std::fabs(dot(lightDir, state.ffnormal)) * lightContrib / lightPdf;

    contrib.visible   = true;
    contrib.lightDir  = lightDir;
    contrib.lightDist = lightDist;
    contrib.radiance  = Li;
  }
  return contrib;
}

// pathtrace.glsl:193-343
inline vec3 PathTrace(PathCtx& c, Ray r)
{
  const Scene& s = *c.scene;
  vec3 radiance = V3(0.0f), throughput = V3(1.0f), absorption = V3(0.0f);
  const int dbg = c.rtx.debugging_mode;

  for(int depth = 0; depth < c.rtx.maxDepth; depth++)
  {
    ClosestHit(c, r);

    if(c.prd.hitT == INFINITY_RT)
    {
      if(dbg != VKRT_DBG_NONE)
      {
        if(depth != c.rtx.maxDepth - 1)
          return V3(0);
        if(dbg == VKRT_DBG_RADIANCE)
          return radiance;
        else if(dbg == VKRT_DBG_WEIGHT)
          return throughput;
        else if(dbg == VKRT_DBG_RAYDIR)
          return (r.direction + V3(1)) * 0.5f;
      }
      vec3 env;
      if(s.sunsky.in_use == 1)
        env = sky::sun_and_sky(s.sunsky, r.direction);
      else
        env = sampleEnv(s, GetSphericalUv(r.direction));
      return radiance + (env * c.rtx.hdrMultiplier * throughput);
    }

    ShadeState sstate = GetShadeState(s, c.prd);

    State state;
    state.position     = sstate.position;
    state.normal       = sstate.normal;
    state.tangent      = sstate.tangent_u;
    state.bitangent    = sstate.tangent_v;
    state.texCoord     = sstate.text_coords;
    state.matID        = sstate.matIndex;
    state.isSubsurface = false;
    state.ffnormal     = dot(state.normal, r.direction) <= 0.0f ? state.normal : -state.normal;

    GetMaterialsAndTextures(s, state, r);
    state.mat.albedo *= sstate.color;

    if(dbg != VKRT_DBG_NONE && dbg < VKRT_DBG_RADIANCE)
      return DebugInfo(c, state);

    if(state.mat.unlit)
      return radiance + state.mat.albedo * throughput;

    if(dot(state.normal, state.ffnormal) > 0.0f)
      absorption = V3(0.0f);

    radiance += state.mat.emission * throughput;
    throughput *= exp3(-absorption * c.prd.hitT);

    VisibilityContribution vcontrib = DirectLight(c, r, state);
    vcontrib.radiance *= throughput;

    vec3  L = V3(0);
    float pdf = 0.f;
    vec3  f = Sample(c, state, -r.direction, state.ffnormal, L, pdf, c.prd.seed);

    if(dot(state.ffnormal, L) < 0.0f)
      absorption = -log3(state.mat.attenuationColor) / V3(state.mat.attenuationDistance);

    if(pdf > 0.0f)
      throughput *= f * std::fabs(dot(state.ffnormal, L)) / pdf;
    else
      break;

    if(dbg != VKRT_DBG_NONE && (depth == c.rtx.maxDepth - 1))
    {
      if(dbg == VKRT_DBG_RADIANCE)
        return vcontrib.radiance;
      else if(dbg == VKRT_DBG_WEIGHT)
        return throughput;
      else if(dbg == VKRT_DBG_RAYDIR)
        return (L + V3(1)) * 0.5f;
    }

    // RR (RR_DEPTH 0)
    float rrPcont = std::fmin(std::fmax(throughput.x, std::fmax(throughput.y, throughput.z)) * state.eta * state.eta + 0.001f, 0.95f);

    r.direction = L;
    r.origin    = OffsetRay(sstate.position, dot(L, state.ffnormal) > 0 ? state.ffnormal : -state.ffnormal);

    if(vcontrib.visible == true)
    {
      Ray  shadowRay{r.origin, vcontrib.lightDir};
      bool inShadow = AnyHit(c, shadowRay, vcontrib.lightDist);
      if(!inShadow)
        radiance += vcontrib.radiance;
    }

    if(rnd(c.prd.seed) >= rrPcont)
      break;
    throughput /= rrPcont;
  }
  return radiance;
}

// pathtrace.glsl:348-387
inline vec3 samplePixel(PathCtx& c, int px, int py, int sx, int sy)
{
  const Scene&       s   = *c.scene;
  const SceneCamera& cam = s.camera;
  vec2 subpixel_jitter;
  if(c.rtx.frame == 0)
    subpixel_jitter = V2(0.5f, 0.5f);
  else
  {
    float jx = rnd(c.prd.seed);
    float jy = rnd(c.prd.seed);
    subpixel_jitter = V2(jx, jy);
  }
  const vec2 pixelCenter = V2((float)px, (float)py) + subpixel_jitter;
  const vec2 inUV        = pixelCenter / V2((float)sx, (float)sy);
  vec2       d           = V2(inUV.x * 2.0f - 1.0f, inUV.y * 2.0f - 1.0f);

  vec4 origin    = mulMV(cam.viewInverse, V4(0, 0, 0, 1));
  vec4 target    = mulMV(cam.projInverse, V4(d.x, d.y, 1, 1));
  vec4 direction = mulMV(cam.viewInverse, V4(normalize(xyz(target)), 0));

  vec3  focalPoint        = cam.focalDist * xyz(direction);
  float cam_r1            = rnd(c.prd.seed) * K_TWO_PI;
  float cam_r2            = rnd(c.prd.seed) * cam.aperture;
  vec4  cam_right         = mulMV(cam.viewInverse, V4(1, 0, 0, 0));
  vec4  cam_up            = mulMV(cam.viewInverse, V4(0, 1, 0, 0));
  vec3  randomAperturePos = (std::cos(cam_r1) * xyz(cam_right) + std::sin(cam_r1) * xyz(cam_up)) * std::sqrt(cam_r2);
  vec3  finalRayDir       = normalize(focalPoint - randomAperturePos);

  Ray  ray{xyz(origin) + randomAperturePos, finalRayDir};
  vec3 radiance = PathTrace(c, ray);

  float lum = dot(radiance, V3(0.212671f, 0.715160f, 0.072169f));
  if(lum > c.rtx.fireflyClampThreshold)
    radiance *= c.rtx.fireflyClampThreshold / lum;
  return radiance;
}

// One pixel of one frame: pathtrace.comp:87-134 / pathtrace.rgen:60-108 (heat-map mode excluded:
// it visualises a hardware clock).  `image` is the RGBA32F accumulation image.
inline void renderPixel(const Scene& s, const RtxState& rtx, int seedMode, int x, int y, float* image, RenderCounters* cnt)
{
  PathCtx c;
  c.scene    = &s;
  c.rtx      = rtx;
  c.seedMode = seedMode;
  c.cnt      = cnt;
  const int W = rtx.size[0], H = rtx.size[1];
  if(seedMode == 0)
    c.prd.seed = tea((uint32_t)(W * y + x), (uint32_t)(rtx.frame * rtx.maxSamples));
  else
    c.prd.seed = initRandom((uint32_t)W, (uint32_t)x, (uint32_t)y, (uint32_t)rtx.frame);

  vec3 pixelColor = V3(0);
  for(int smpl = 0; smpl < rtx.maxSamples; ++smpl)
    pixelColor += samplePixel(c, x, y, W, H);
  pixelColor /= (float)rtx.maxSamples;

  float* px = image + (size_t(y) * W + x) * 4;
  if(rtx.frame > 0)
  {
    vec3  old_color = V3(px[0], px[1], px[2]);
    vec3  nr        = mix(old_color, pixelColor, 1.0f / float(rtx.frame + 1));
    px[0] = nr.x; px[1] = nr.y; px[2] = nr.z; px[3] = 1.f;
  }
  else
  {
    px[0] = pixelColor.x; px[1] = pixelColor.y; px[2] = pixelColor.z; px[3] = 1.f;
  }
}

}  // namespace orc
